"""Layout adapter between the reference pipeline and the solvers (row N4 of SURVEY.md 8(f)).

The reference's drivers hold voxel matrices time-major, ``[T, V]`` -- what
``NiftiMasker.fit_transform`` returns (examples/icassp_2019/validation.py:90-93) -- and iterate over
``voxels.T`` (validation.py:43-47, simulation.py:64-67).  The batched solvers want ``[V, T]`` with T
contiguous.  Both directions run as one tile-movement kernel on the device (TMA, csrc/pb_transpose_tma.cuh);
large host matrices are uploaded and transposed chunk by chunk on two streams.
"""
from __future__ import annotations

import numpy as np
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


_STREAM_MIN_BYTES = 64 << 20      # host matrices above this are uploaded and transposed chunk by chunk
_STREAM_CHUNK_VOXELS = 1 << 16


def _from_host_streamed(a, dtype, chunk):
    """Host ``[T, V]`` -> device ``[V, T]`` in column chunks: while chunk i is transposed on the device, chunk
    i + 1 is staged in pinned memory and uploaded on the other stream (the matrix is never resident twice)."""
    T, V = a.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((V, T), dtype=dtype, device=dev)
    fn = _lib.fn("pb_transpose", dtype)
    main = torch.cuda.current_stream()
    slots = [{"stream": torch.cuda.Stream(device=dev), "pin": torch.empty((T, chunk), dtype=dtype, pin_memory=True),
              "dev": torch.empty((T, chunk), dtype=dtype, device=dev), "done": None} for _ in range(2)]
    for i, v0 in enumerate(range(0, V, chunk)):
        v1 = min(v0 + chunk, V)
        n = v1 - v0
        slot = slots[i % 2]
        if slot["done"] is not None:
            slot["done"].synchronize()              # the pinned buffer is free again
        pin = slot["pin"][:, :n] if n == chunk else slot["pin"].reshape(-1)[:T * n].reshape(T, n)
        pin.copy_(a[:, v0:v1])                      # strided host gather into a dense [T, n] block
        d = slot["dev"] if n == chunk else slot["dev"].reshape(-1)[:T * n].reshape(T, n)
        st = slot["stream"]
        st.wait_stream(main)
        with torch.cuda.stream(st):
            d.copy_(pin, non_blocking=True)
            rc = fn(ptr(d), ptr(out[v0:v1]), T, n, st.cuda_stream)
            _lib.check(rc, "pb_transpose")
            slot["done"] = torch.cuda.Event()
            slot["done"].record(st)
    for slot in slots:
        main.wait_stream(slot["stream"])
    return out


def voxels_from_timeseries(voxels_tv, chunk_voxels=None):
    """``[T, V]`` (time-major, as the reference's drivers hold it) -> CUDA ``[V, T]`` for ``bd`` / ``deconv``.

    A HOST matrix above 64 MB is streamed: column chunks of ``chunk_voxels`` voxels go through pinned memory,
    upload and transposition of consecutive chunks overlap on two streams."""
    is_host = not (isinstance(voxels_tv, torch.Tensor) and voxels_tv.is_cuda)
    if is_host:
        dtype = pick_dtype(voxels_tv)
        a = voxels_tv if isinstance(voxels_tv, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(voxels_tv))
        if a.dim() != 2:
            raise ValueError("expected a 2-D matrix")
        nbytes = a.numel() * a.element_size()
        if chunk_voxels is not None or nbytes > _STREAM_MIN_BYTES:
            from ._array import require_cuda
            require_cuda()
            chunk = int(chunk_voxels or _STREAM_CHUNK_VOXELS)
            if a.shape[1] > chunk:
                return _from_host_streamed(a.to(dtype), dtype, chunk)
    return _transpose(voxels_tv)


def timeseries_from_voxels(signals_vt):
    """``[V, T]`` solver output -> CUDA ``[T, V]`` (the layout ``NiftiMasker.inverse_transform`` expects)."""
    return _transpose(signals_vt)
