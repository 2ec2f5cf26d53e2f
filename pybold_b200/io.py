"""Layout adapter between the reference pipeline and the solvers (row N4 of SURVEY.md 8(f)).

The reference's drivers hold voxel matrices time-major, ``[T, V]`` -- what
``NiftiMasker.fit_transform`` returns (examples/icassp_2019/validation.py:90-93) -- and iterate over
``voxels.T`` (validation.py:43-47, simulation.py:64-67).  The batched solvers want ``[V, T]`` with T
contiguous.  Both directions run as one tiled transpose kernel on the device.
"""
from __future__ import annotations

import torch

from . import _lib
from ._array import pick_dtype, ptr, stream_ptr, to_device


def _transpose(a):
    dtype = pick_dtype(a)
    ad = to_device(a, dtype)
    if ad.dim() != 2:
        raise ValueError("expected a 2-D matrix")
    rows, cols = ad.shape
    out = torch.empty((cols, rows), dtype=dtype, device=ad.device)
    rc = _lib.fn("pb_transpose", dtype)(ptr(ad), ptr(out), rows, cols, stream_ptr())
    _lib.check(rc, "pb_transpose")
    return out


def voxels_from_timeseries(voxels_tv):
    """``[T, V]`` (time-major, as the reference's drivers hold it) -> CUDA ``[V, T]`` for ``bd`` / ``deconv``."""
    return _transpose(voxels_tv)


def timeseries_from_voxels(signals_vt):
    """``[V, T]`` solver output -> CUDA ``[T, V]`` (the layout ``NiftiMasker.inverse_transform`` expects)."""
    return _transpose(signals_vt)
