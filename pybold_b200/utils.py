"""Hot-path subset of ``pybold/utils.py``: the power-iteration Lipschitz estimate."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._array import pick_dtype, ptr, stream_ptr, to_device
from .linear import ConvAndLinear, DiscretInteg


def spectral_radius_est(L, x_shape, nb_iter=30, tol=1.0e-6, verbose=False, x0=None):
    """Spectral radius of ``L.adj o L.op`` by power iteration (pybold/utils.py:94-109).

    Like the reference, the start vector is drawn from NumPy's *global* generator
    (``np.random.randn(*x_shape)``) unless ``x0`` is given, so seeding ``np.random`` the same
    way gives the same estimate.  ``L`` must be ``ConvAndLinear(DiscretInteg(), kernel, T)``:
    the whole iteration then runs in one kernel.  ``x0`` / the kernel may be batched ``[V, .]``.
    """
    if not (isinstance(L, ConvAndLinear) and isinstance(L.M, DiscretInteg)):
        raise NotImplementedError("spectral_radius_est runs on ConvAndLinear(DiscretInteg(), ...) only")
    if x0 is None:
        x0 = np.random.randn(*x_shape)
    dtype = pick_dtype(L.k, x0)
    xd = to_device(x0, dtype)
    kd = to_device(L.k, dtype)
    scalar = xd.dim() == 1 and kd.dim() == 1
    x2 = xd.reshape(1, -1) if xd.dim() == 1 else xd
    k2 = kd.reshape(1, -1) if kd.dim() == 1 else kd
    V = max(x2.shape[0], k2.shape[0])
    T, K = x2.shape[1], k2.shape[1]
    out = torch.empty(V, dtype=dtype, device=xd.device)
    rc = _lib.fn("pb_lipschitz_power", dtype)(
        ptr(k2), K if k2.shape[0] > 1 else 0, ptr(x2), T if x2.shape[0] > 1 else 0,
        int(nb_iter), float(tol), ptr(out), V, T, K, stream_ptr())
    _lib.check(rc, "pb_lipschitz_power")
    if scalar:
        return float(out[0])
    return out if isinstance(x0, torch.Tensor) else out.cpu().numpy()
