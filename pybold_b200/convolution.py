"""Short-kernel convolution / correlation -- device mirror of ``pybold/convolution.py``.

The reference has an FFT path with custom padding (``spectral_*``) and O(T K) loops
(``simple_*``) computing the same thing (its tests pin them to each other at 1e-7,
pybold/tests/test_convolution.py).  Here all four names run the direct K-tap kernel.
Inputs may be 1-D (one voxel) or ``[V, T]``; the kernel ``k`` may be ``[K]`` or ``[V, K]``.
"""
from __future__ import annotations

import torch

from . import _lib
from ._array import like_input, pick_dtype, ptr, stream_ptr, to_device


def _run(name, k, x):
    dtype = pick_dtype(k, x)
    xd = to_device(x, dtype)
    kd = to_device(k, dtype)
    one_d = xd.dim() == 1
    x2 = xd.reshape(1, -1) if one_d else xd
    V, T = x2.shape
    if kd.dim() == 1:
        K, stride = kd.numel(), 0
    else:
        if kd.shape[0] != V:
            raise ValueError("per-voxel kernels need shape [V, K]")
        K, stride = kd.shape[1], kd.shape[1]
    out = torch.empty_like(x2)
    rc = _lib.fn(name, dtype)(ptr(kd), stride, ptr(x2), ptr(out), V, T, K, stream_ptr())
    _lib.check(rc, name)
    out = out.reshape(-1) if one_d else out
    return like_input(out, x)


def simple_convolve(k, x, dim_out=None):
    """out[i] = sum_j k[j] x[i-j] (pybold/convolution.py:135-164)."""
    if dim_out is not None and dim_out != (x.shape[-1]):
        raise NotImplementedError("only square (dim_out == len(x)) convolutions are on the hot path")
    return _run("pb_conv_op", k, x)


def simple_retro_convolve(k, x, dim_out=None):
    """out[i] = sum_j k[j] x[i+j] (pybold/convolution.py:167-196)."""
    if dim_out is not None and dim_out != (x.shape[-1]):
        raise NotImplementedError("only square (dim_out == len(x)) convolutions are on the hot path")
    return _run("pb_conv_adj", k, x)


def spectral_convolve(k, x):
    """Same result as the reference's padded FFT convolution (pybold/convolution.py:9-30)."""
    return _run("pb_conv_op", k, x)


def spectral_retro_convolve(k, x):
    """Same result as pybold/convolution.py:33-54."""
    return _run("pb_conv_adj", k, x)
