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


def _resize_last(x, n):
    """x truncated or zero-padded along its last axis to n samples (host or device array)."""
    T = x.shape[-1]
    if n == T:
        return x
    if n < T:
        return x[..., :n]
    if isinstance(x, torch.Tensor):
        return torch.nn.functional.pad(x, (0, n - T))
    import numpy as np
    pad = [(0, 0)] * (np.ndim(x) - 1) + [(0, n - T)]
    return np.pad(np.asarray(x), pad)


def simple_convolve(k, x, dim_out=None):
    """out[i] = sum_j k[j] x[i-j], i < dim_out (pybold/convolution.py:135-164); dim_out defaults to len(x)."""
    T = x.shape[-1]
    if dim_out is None or dim_out == T:
        return _run("pb_conv_op", k, x)
    if dim_out > T:                      # outputs past the input: x continues with zeros
        return _run("pb_conv_op", k, _resize_last(x, dim_out))
    return _resize_last(_run("pb_conv_op", k, x), dim_out)


def simple_retro_convolve(k, x, dim_out=None):
    """out[i] = sum_j k[j] x[i+j], i < dim_out (pybold/convolution.py:167-196); dim_out defaults to len(x)."""
    T = x.shape[-1]
    if dim_out is None or dim_out == T:
        return _run("pb_conv_adj", k, x)
    if dim_out > T:
        return _run("pb_conv_adj", k, _resize_last(x, dim_out))
    return _resize_last(_run("pb_conv_adj", k, x), dim_out)


def toeplitz_from_kernel(k, dim_in, dim_out=None):
    """Dense ``[dim_out, dim_in]`` Toeplitz matrix of ``k.conv(.)`` (pybold/convolution.py:105-132):
    ``K[i, c] = k[i - c]``.  The solvers never build it; provided for callers of the reference module."""
    if dim_out is None:
        dim_out = dim_in
    dtype = pick_dtype(k)
    kd = to_device(k, dtype).reshape(-1)
    out = torch.empty((int(dim_out), int(dim_in)), dtype=dtype, device=kd.device)
    rc = _lib.fn("pb_toeplitz", dtype)(ptr(kd), kd.numel(), ptr(out), int(dim_out), int(dim_in), stream_ptr())
    _lib.check(rc, "pb_toeplitz")
    return like_input(out, k)


def spectral_convolve(k, x):
    """Same result as the reference's padded FFT convolution (pybold/convolution.py:9-30)."""
    return _run("pb_conv_op", k, x)


def spectral_retro_convolve(k, x):
    """Same result as pybold/convolution.py:33-54."""
    return _run("pb_conv_adj", k, x)
