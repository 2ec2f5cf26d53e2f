"""Hot-path subset of ``pybold/utils.py``: the power-iteration Lipschitz estimate, ``inf_norm`` and the
relative-error metric of the ICASSP-2019 simulation (row N3)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._array import pick_dtype, ptr, stream_ptr, to_device
from .linear import ConvAndLinear, DiscretInteg
from .noise import mad, mad_daub_noise_est  # noqa: F401  (same module as in the reference, pybold/utils.py:10-25)


def spectral_radius_est(L, x_shape, nb_iter=30, tol=1.0e-6, verbose=False, x0=None):
    """Spectral radius of ``L.adj o L.op`` by power iteration (pybold/utils.py:94-109).

    Like the reference, the start vector is drawn from NumPy's *global* generator
    (``np.random.randn(*x_shape)``) unless ``x0`` is given, so seeding ``np.random`` the same
    way gives the same estimate.  For ``L = ConvAndLinear(DiscretInteg(), kernel, T)`` -- what ``deconv``
    builds -- the whole iteration runs in ONE kernel (``x0`` / the kernel may be batched ``[V, .]``); any
    other object with ``op`` / ``adj`` is iterated call by call like the reference does.
    """
    if x0 is None:
        x0 = np.random.randn(*x_shape)
    if not (isinstance(L, ConvAndLinear) and isinstance(L.M, DiscretInteg)):
        return _power_iteration_generic(L, x0, nb_iter, tol, verbose)
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


def _power_iteration_generic(L, x0, nb_iter, tol, verbose):
    """pybold/utils.py:94-109 for ANY object with ``op`` / ``adj`` (the reference's duck-typed operator
    protocol): the loop of the reference, statement by statement, around the operator's own calls.  With the
    operators of ``pybold_b200.linear`` every ``op`` / ``adj`` is a device kernel; the two norms per step are
    the only arithmetic done here."""
    def norm(a):
        if isinstance(a, torch.Tensor):
            return float(torch.linalg.vector_norm(a.double()))
        return float(np.linalg.norm(np.asarray(a, dtype=np.float64)))

    x_old = x0
    x_new = x_old
    for i in range(nb_iter):
        x_new = L.adj(L.op(x_old)) / norm(x_old)
        if abs(norm(x_new) - norm(x_old)) < tol:
            if verbose:
                print("Spectral radius estimation converged at iteration {0}".format(i))
            break
        x_old = x_new
    return norm(x_new)


def _rows(x):
    dtype = pick_dtype(x)
    xd = to_device(x, dtype)
    one_d = xd.dim() == 1
    x2 = (xd.reshape(1, -1) if one_d else xd).contiguous()
    return dtype, xd, one_d, x2


def inf_norm(arrays, axis=1):
    """Inf-norm normalisation ``x / (max|x| + 1e-12)`` (pybold/utils.py:112-138): 1-D arrays as a whole,
    2-D arrays row by row (``axis=1``; ``axis=0`` normalises the columns), lists element-wise."""
    from ._array import like_input
    if isinstance(arrays, list):
        return [inf_norm(a, axis=axis) for a in arrays]
    nd = arrays.dim() if isinstance(arrays, torch.Tensor) else np.ndim(arrays)
    if nd == 3:                                   # normalised as a whole, like 1-D (utils.py:131-132)
        dtype = pick_dtype(arrays)
        xd = to_device(arrays, dtype)
        flat = inf_norm(xd.reshape(-1))
        return like_input(flat.reshape(xd.shape), arrays)
    if nd not in (1, 2):
        raise ValueError("inf-norm normalization only handle 1D, 2D or 3D arrays")
    dtype, xd, one_d, x2 = _rows(arrays)
    if not one_d and axis == 0:
        x2 = x2.t().contiguous()
    out = torch.empty_like(x2)
    rc = _lib.fn("pb_inf_norm", dtype)(ptr(x2), ptr(out), x2.shape[0], x2.shape[1], stream_ptr())
    _lib.check(rc, "pb_inf_norm")
    if not one_d and axis == 0:
        out = out.t().contiguous()
    return like_input(out.reshape(-1) if one_d else out, arrays)


def rel_l2_err(est, ref):
    """Per-voxel relative L2 error ``||est_v - ref_v|| / ||ref_v||`` of the ICASSP-2019 simulation
    (examples/icassp_2019/simulation.py:143-147); ``ref`` is ``[V, T]`` or one shared row ``[T]``."""
    from ._array import like_input
    dtype = pick_dtype(est, ref)
    e2 = to_device(est, dtype)
    r2 = to_device(ref, dtype).contiguous()
    e2 = (e2.reshape(1, -1) if e2.dim() == 1 else e2).contiguous()
    V, T = e2.shape
    if r2.shape[-1] != T or (r2.dim() == 2 and r2.shape[0] not in (1, V)):
        raise ValueError("rel_l2_err: est %s and ref %s do not match" % (tuple(e2.shape), tuple(r2.shape)))
    stride = T if (r2.dim() == 2 and r2.shape[0] == V and V > 1) else 0
    out = torch.empty(V, dtype=dtype, device=e2.device)
    rc = _lib.fn("pb_rel_l2_err", dtype)(ptr(e2), ptr(r2), stride, ptr(out), V, T, stream_ptr())
    _lib.check(rc, "pb_rel_l2_err")
    return like_input(out, est)


class Tracker:
    """Iterate callback that records ``f(x, *args)`` (pybold/utils.py:27-45), for user code that still
    drives a SciPy optimiser; the device theta solver does not go through iterate callbacks."""

    def __init__(self, f, args, verbose=0):
        self.f, self.args, self.verbose = f, list(args), verbose
        self.J, self.idx = [], 0

    def __call__(self, x):
        self.idx += 1
        value = self.f(x, *self.args)
        if self.verbose > 2:
            print("At iterate {0}, tracked function = {1:.6f}".format(self.idx, value))
        self.J.append(value)
