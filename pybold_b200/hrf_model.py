"""SPM HRF with a time-dilation parameter -- device mirror of ``pybold/hrf_model.py``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._array import like_input, pick_dtype, ptr, require_cuda, stream_ptr

MIN_DELTA = 0.5   # pybold/hrf_model.py:8
MAX_DELTA = 2.0   # pybold/hrf_model.py:9


def hrf_len(t_r, dur):
    """Number of taps ``spm_hrf`` returns for this ``t_r`` / ``dur`` (hrf_model.py:36)."""
    K = _lib.lib.pb_hrf_len(float(t_r), float(dur))
    _lib.check(min(K, 0), "pb_hrf_len")
    return K


_DEFAULT_SHAPE = (0.001, 6, 16.0, 1.0, 1.0, 0.167, 0.0)


def spm_hrf(delta, t_r=1.0, dur=60.0, normalized_hrf=True, dt=0.001, p_delay=6,
            undershoot=16.0, p_disp=1.0, u_disp=1.0, p_u_ratio=0.167, onset=0.0):
    """Same signature and return as ``pybold.hrf_model.spm_hrf`` (hrf_model.py:12-39).

    ``delta`` may also be a length-V array / tensor: the result is then ``[V, K]``.  The
    reference's default shape parameters run on the integer-power kernel the solvers use
    (``pb_spm_hrf_*``), any other combination on ``pb_spm_hrf_ex_*``.  Like the reference the
    time axis is shifted by ``onset / dt`` (hrf_model.py:25), i.e. ``onset=0.004`` means 4 s.
    """
    import ctypes
    shape = (dt, p_delay, undershoot, p_disp, u_disp, p_u_ratio, onset)
    require_cuda()
    scalar = not isinstance(delta, torch.Tensor) and np.ndim(delta) == 0
    dtype = pick_dtype(delta) if not scalar else torch.float64
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(delta, torch.Tensor):
        th = delta.to(device=dev, dtype=dtype).reshape(-1).contiguous()
    else:
        th = torch.as_tensor(np.asarray(delta, dtype=np.float64).reshape(-1)).to(device=dev, dtype=dtype)
    lo, hi = float(th.min()), float(th.max())
    if lo < MIN_DELTA or hi > MAX_DELTA or lo != lo:                 # hrf_model.py:17-21
        raise ValueError("delta should belong in [{0}, {1}]; wich correspond to a max FWHM of "
                         "10.52s and a min FWHM of 2.80s, got delta = {2}".format(
                             MIN_DELTA, MAX_DELTA, lo if lo < MIN_DELTA else hi))
    V = th.numel()
    if tuple(float(v) for v in shape) == tuple(float(v) for v in _DEFAULT_SHAPE):
        K = hrf_len(t_r, dur)
        out = torch.empty((V, K), dtype=dtype, device=dev)
        rc = _lib.fn("pb_spm_hrf", dtype)(ptr(th), float(t_r), float(dur), int(bool(normalized_hrf)),
                                          ptr(out), V, K, stream_ptr())
        _lib.check(rc, "pb_spm_hrf")
    else:
        K = _lib.lib.pb_hrf_len_ex(float(t_r), float(dur), float(dt))
        _lib.check(min(K, 0), "pb_hrf_len_ex")
        out = torch.empty((V, K), dtype=dtype, device=dev)
        shape7 = (ctypes.c_double * 7)(*[float(v) for v in shape])
        rc = _lib.fn("pb_spm_hrf_ex", dtype)(ptr(th), float(t_r), float(dur), int(bool(normalized_hrf)),
                                             shape7, ptr(out), V, K, stream_ptr())
        _lib.check(rc, "pb_spm_hrf_ex")
    n_fine = int(float(dur) / dt)
    t_hrf = (np.linspace(0, dur, n_fine) - float(onset) / dt)[::int(t_r / dt)]
    if scalar:
        return out[0].cpu().numpy(), t_hrf
    return like_input(out, delta), t_hrf
