"""Array plumbing between NumPy / torch inputs and the device buffers the C ABI works on.

PyTorch is used only to own device memory and streams.  There is deliberately no code
path here that computes anything on the CPU.
"""
from __future__ import annotations

import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pybold_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")


def pick_dtype(*arrays, dtype=None):
    """float32 only when every floating input is float32; the reference is float64 throughout."""
    if dtype is not None:
        return {np.float32: torch.float32, np.float64: torch.float64,
                "float32": torch.float32, "float64": torch.float64}.get(dtype, dtype)
    seen = []
    for a in arrays:
        if isinstance(a, torch.Tensor):
            seen.append(a.dtype)
        elif isinstance(a, np.ndarray) and a.dtype.kind == "f":
            seen.append(torch.float32 if a.dtype == np.float32 else torch.float64)
    if seen and all(d == torch.float32 for d in seen):
        return torch.float32
    return torch.float64


def to_device(x, dtype, device=None):
    """Contiguous CUDA tensor of ``dtype`` (copies host data through pinned memory when large)."""
    require_cuda()
    device = device or torch.device("cuda", torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(x))
    return upload(arr, dtype, device)


_PINNED_UPLOAD_LIMIT = 16 << 20


def upload(arr, dtype, device):
    """NumPy array -> device tensor of ``dtype``.  Small and medium arrays go through a pinned block (PyTorch
    caches them) and an asynchronous copy: a pageable upload makes the host wait on the launch stream, and
    that wait was measured to stall for tens of milliseconds now and then with the device idle
    (tools/debug_e2e.py).  Larger arrays take the driver's own staged path."""
    device = torch.device(device)
    src = torch.from_numpy(arr)
    if device.type != "cuda" or arr.nbytes > _PINNED_UPLOAD_LIMIT or arr.size == 0:
        return src.to(device=device, dtype=dtype).contiguous()
    stage = torch.empty(src.shape, dtype=dtype, pin_memory=True)
    stage.copy_(src)                                   # converts on the host
    return stage.to(device=device, non_blocking=True)


_scalar_cache = {}


def _device_scalar(v, dtype, device):
    """One-element device tensor holding ``v``; cached (read-only by convention: kernels only load it), so a
    scalar lbda / theta_0 costs no upload, and no host wait on the stream, after its first use."""
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (v, dtype, device)
    t = _scalar_cache.get(key)
    if t is None:
        if len(_scalar_cache) >= 256:
            _scalar_cache.clear()
        t = torch.full((1,), v, dtype=dtype, device=device)
        if device.type == "cuda":
            torch.cuda.current_stream(device).synchronize()      # complete before another stream may read it
        _scalar_cache[key] = t
    return t


def per_voxel(value, V, dtype, device, name):
    """Scalar or length-V parameter -> (tensor, element stride) as the C ABI wants it."""
    if isinstance(value, torch.Tensor):
        t = value.to(device=device, dtype=dtype).reshape(-1).contiguous()
    else:
        arr = np.asarray(value, dtype=np.float64).reshape(-1)
        if arr.size == 1:
            return _device_scalar(float(arr[0]), dtype, torch.device(device)), 0
        t = upload(arr, dtype, device)
    if t.numel() == 1:
        return t, 0
    if t.numel() != V:
        raise ValueError("%s must be a scalar or have one entry per voxel (%d), got %d"
                         % (name, V, t.numel()))
    return t, 1


def like_input(t, template):
    """Return ``t`` as the kind of array the caller passed in (NumPy in -> NumPy out)."""
    if isinstance(template, torch.Tensor):
        if template.is_cuda:
            return t
        return t.cpu()          # host tensor in -> host tensor out
    return t.detach().cpu().numpy()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()
