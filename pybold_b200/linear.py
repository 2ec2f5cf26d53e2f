"""Linear operators with the reference's ``op`` / ``adj`` protocol (``pybold/linear.py``)."""
from __future__ import annotations

import torch

from . import _lib
from ._array import like_input, pick_dtype, ptr, stream_ptr, to_device
from .convolution import _run as _conv_run


class DiscretInteg:
    """Integrator: ``op`` = running sum, ``adj`` = reversed running sum (linear.py:9-43)."""

    def __init__(self):
        pass

    @staticmethod
    def _run(name, x):
        dtype = pick_dtype(x)
        xd = to_device(x, dtype)
        one_d = xd.dim() == 1
        x2 = xd.reshape(1, -1) if one_d else xd
        out = torch.empty_like(x2)
        rc = _lib.fn(name, dtype)(ptr(x2), ptr(out), x2.shape[0], x2.shape[1], stream_ptr())
        _lib.check(rc, name)
        return like_input(out.reshape(-1) if one_d else out, x)

    def op(self, x):
        return self._run("pb_integ_op", x)

    def adj(self, x):
        return self._run("pb_integ_adj", x)


class ConvAndLinear:
    """Linear operator followed by a convolution (linear.py:46-113).

    ``op(x) = k * M.op(x)``, ``adj(x) = M.adj(k^T * x)``.  With ``M = DiscretInteg()`` (the only
    combination the solvers use) both run as one fused kernel.  ``spectral_conv`` is accepted
    for signature compatibility; both settings give the direct K-tap result.
    """

    def __init__(self, M, kernel, dim_in, dim_out=None, spectral_conv=False):
        self.M = M
        self.k = kernel
        self.dim_in = dim_in
        self.dim_out = dim_in if dim_out is None else dim_out
        self.spectral_conv = spectral_conv

    def op(self, x):
        if self.dim_out != self.dim_in:          # rectangular Toeplitz (linear.py:69): not on the hot path
            from .convolution import simple_convolve
            return simple_convolve(self.k, self.M.op(x), self.dim_out)
        if isinstance(self.M, DiscretInteg):
            return _conv_run("pb_hrfinteg_op", self.k, x)
        return _conv_run("pb_conv_op", self.k, self.M.op(x))

    def adj(self, x):
        if self.dim_out != self.dim_in:
            from .convolution import simple_retro_convolve
            return self.M.adj(simple_retro_convolve(self.k, x, self.dim_in))
        if isinstance(self.M, DiscretInteg):
            return _conv_run("pb_hrfinteg_adj", self.k, x)
        return self.M.adj(_conv_run("pb_conv_adj", self.k, x))
