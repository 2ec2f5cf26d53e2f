"""Synthetic BOLD voxels for tests and benchmarks.

Own generator (the reference's ``pybold/data.py`` is broken on NumPy >= 1.24 and is out
of scope): per voxel ``v`` a block paradigm of ``nb_events`` unit boxcars of
``ceil(avg_dur / t_r)`` samples (ICASSP-2019 setup, examples/icassp_2019/simulation.py:105-111),
convolved with a normalised SPM HRF of dilation ``delta_v ~ U[0.7, 1.3]`` and Gaussian noise
scaled to ``snr_db`` with the rule of pybold/data.py:439-444.  Everything is float64 and
depends only on ``seed0 + v`` so that any voxel range can be regenerated on any rank.
"""
from __future__ import annotations

import math

import numpy as np


def _hrf_taps(delta, t_r, dur, dt=0.001):
    """Closed-form SPM taps at the kept samples, non-normalised (pybold/hrf_model.py:12-39)."""
    n_fine = int(float(dur) / dt)
    stride = int(t_r / dt)
    m = np.arange(0, n_fine, stride, dtype=np.float64)
    s = float(delta) * (m * (float(dur) / (n_fine - 1))) - dt
    pos = s > 0
    sp = np.where(pos, s, 1.0)
    e = np.exp(-sp)
    h = sp ** 5 * e / math.factorial(5) - 0.167 * sp ** 15 * e / math.factorial(15)
    return np.where(pos, h, 0.0)


def gen_voxels(n_voxels, n_scans, t_r=1.0, hrf_dur=20.0, snr_db=10.0, nb_events=5,
               avg_dur=12.0, seed0=0, first_voxel=0, delta_range=(0.7, 1.3),
               return_truth=False):
    """Return ``y[V, T]`` float64 (and ``z_true[V, T]``, ``delta[V]`` when asked)."""
    T = int(n_scans)
    blk = int(math.ceil(avg_dur / t_r))
    y = np.empty((n_voxels, T))
    z_all = np.empty((n_voxels, T)) if return_truth else None
    deltas = np.empty(n_voxels)
    hi = max(T - blk - 1, 1)
    for i in range(n_voxels):
        rng = np.random.RandomState(seed0 + first_voxel + i)
        onsets = rng.randint(0, hi, nb_events)
        delta = rng.uniform(*delta_range)
        z = np.zeros(T)
        for o in onsets:
            z[o:o + blk] += 1.0
        h = _hrf_taps(delta, t_r, hrf_dur)
        h = h / np.max(np.abs(h))
        x = np.convolve(h, z)[:T]
        n = rng.randn(T)
        n *= np.linalg.norm(x) / (np.linalg.norm(n) + np.finfo(float).eps) / 10.0 ** (snr_db / 20.0)
        y[i] = x + n
        deltas[i] = delta
        if return_truth:
            z_all[i] = z
    if return_truth:
        return y, z_all, deltas
    return y


def gen_voxels_chunked(n_voxels, n_scans, t_r=1.0, hrf_dur=20.0, snr_db=10.0, nb_events=5,
                       avg_dur=12.0, seed0=0, first_voxel=0, delta_range=(0.7, 1.3), chunk=1024,
                       dtype=np.float32):
    """Vectorised variant for benchmark-sized batches (100 k+ voxels).

    Same model as :func:`gen_voxels`, but random numbers are drawn per aligned chunk of
    ``chunk`` voxels (``RandomState(seed0 + chunk_index)``), so any voxel range can still be
    regenerated independently on any rank while the generation runs at NumPy speed.
    """
    T = int(n_scans)
    blk = int(math.ceil(avg_dur / t_r))
    hi = max(T - blk - 1, 1)
    out = np.empty((n_voxels, T), dtype=dtype)
    n_fine = int(float(hrf_dur) / 0.001)
    stride = int(t_r / 0.001)
    tm = np.arange(0, n_fine, stride, dtype=np.float64) * (float(hrf_dur) / (n_fine - 1))
    K = len(tm)
    first_chunk = first_voxel // chunk
    last_chunk = (first_voxel + n_voxels - 1) // chunk if n_voxels else first_chunk - 1
    for c in range(first_chunk, last_chunk + 1):
        rng = np.random.RandomState(seed0 + c)
        onsets = rng.randint(0, hi, (chunk, nb_events))
        delta = rng.uniform(delta_range[0], delta_range[1], chunk)
        noise = rng.randn(chunk, T)
        # block paradigm: +1 at onset, -1 at onset + blk, integrated
        dz = np.zeros((chunk, T + blk + 1))
        rows = np.repeat(np.arange(chunk), nb_events)
        np.add.at(dz, (rows, onsets.ravel()), 1.0)
        np.add.at(dz, (rows, onsets.ravel() + blk), -1.0)
        z = np.cumsum(dz[:, :T], axis=1)
        s = delta[:, None] * tm[None, :] - 0.001
        pos = s > 0
        sp = np.where(pos, s, 1.0)
        e = np.exp(-sp)
        h = np.where(pos, sp ** 5 * e / 120.0 - 0.167 * sp ** 15 * e / 1307674368000.0, 0.0)
        h /= np.max(np.abs(h), axis=1, keepdims=True)
        x = np.zeros((chunk, T))
        for j in range(min(K, T)):
            x[:, j:] += h[:, j:j + 1] * z[:, :T - j]
        scale = (np.linalg.norm(x, axis=1) / (np.linalg.norm(noise, axis=1) + np.finfo(float).eps)
                 / 10.0 ** (snr_db / 20.0))
        yc = x + noise * scale[:, None]
        g0 = c * chunk
        lo = max(first_voxel, g0)
        hi_v = min(first_voxel + n_voxels, g0 + chunk)
        out[lo - first_voxel:hi_v - first_voxel] = yc[lo - g0:hi_v - g0]
    return out


# ------------------------------------------------------------------------------------------------
# "philox-v1": the counter-based generator of the on-device batch generator (row N3).
# `gen_voxels_philox` is its NumPy statement (the specification the CUDA kernel in
# csrc/pb_synth.cuh is tested against); `gen_voxels_device` runs the kernel.
# ------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK, _S32 = np.uint64(0xFFFFFFFF), np.uint64(32)


def philox4x32(ctr, key, rounds=10):
    """Philox4x32-10 (Salmon et al., SC'11): ``ctr`` = 4 and ``key`` = 2 broadcastable arrays of
    32-bit words; returns the four output words as uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & _MASK for x in ctr]
    k = [np.asarray(x, dtype=np.uint64) & _MASK for x in key]
    for r in range(rounds):
        if r:
            k = [(k[0] + _W0) & _MASK, (k[1] + _W1) & _MASK]
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [(p1 >> _S32) ^ c[1] ^ k[0], p1 & _MASK, (p0 >> _S32) ^ c[3] ^ k[1], p0 & _MASK]
    return [x.astype(np.uint32) for x in c]


def _u01(u):
    return (u.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def gen_voxels_philox(n_voxels, n_scans, t_r=1.0, hrf_dur=20.0, snr_db=10.0, nb_events=5,
                      avg_dur=12.0, seed=0, first_voxel=0, delta_range=(0.7, 1.3)):
    """NumPy statement of "philox-v1" (see csrc/pb_synth.cuh).  Returns ``(y, z, delta)`` float64."""
    V, T = int(n_voxels), int(n_scans)
    blk = int(math.ceil(avg_dur / t_r))
    n_onset = max(T - blk - 1, 1)
    g = np.arange(first_voxel, first_voxel + V, dtype=np.uint64)
    g_lo, g_hi = g & _MASK, g >> _S32
    key = (np.uint64(seed) & _MASK, np.uint64(seed) >> _S32)
    nq = (nb_events + 1 + 3) // 4
    u = np.concatenate([np.stack(philox4x32((np.uint64(q), 0, g_lo, g_hi), key), axis=1)
                        for q in range(nq)], axis=1)                    # [V, 4 nq]
    delta = delta_range[0] + (delta_range[1] - delta_range[0]) * _u01(u[:, 0])
    onsets = ((u[:, 1:1 + nb_events].astype(np.uint64) * np.uint64(n_onset)) >> _S32).astype(np.int64)
    idx = np.arange(T)[None, None, :]
    z = np.sum((idx >= onsets[:, :, None]) & (idx < onsets[:, :, None] + blk), axis=1).astype(np.float64)
    n_fine = int(float(hrf_dur) / 0.001)
    stride = int(t_r / 0.001)
    tm = np.arange(0, n_fine, stride, dtype=np.float64) * (float(hrf_dur) / (n_fine - 1))
    s = delta[:, None] * tm[None, :] - 0.001
    pos = s > 0
    sp = np.where(pos, s, 1.0)
    s5 = sp ** 5
    h = np.where(pos, np.exp(-sp) * s5 * (1.0 / 120.0 - 0.167 * (s5 * s5) * (1.0 / 1307674368000.0)), 0.0)
    h /= np.max(np.abs(h), axis=1, keepdims=True)
    x = np.zeros((V, T))
    for j in range(min(h.shape[1], T)):
        x[:, j:] += h[:, j:j + 1] * z[:, :T - j]
    ngrp = (T + 3) // 4
    pgrid = np.arange(ngrp, dtype=np.uint64)[None, :]
    r = philox4x32((pgrid, 1, g_lo[:, None], g_hi[:, None]), key)       # 4 x [V, ngrp]
    noise = np.empty((V, ngrp, 4))
    for hf in range(2):
        rad = np.sqrt(-2.0 * np.log(_u01(r[2 * hf])))
        ang = 2.0 * np.pi * _u01(r[2 * hf + 1])
        noise[:, :, 2 * hf] = rad * np.cos(ang)
        noise[:, :, 2 * hf + 1] = rad * np.sin(ang)
    noise = noise.reshape(V, 4 * ngrp)[:, :T]
    scale = (np.linalg.norm(x, axis=1) / (np.linalg.norm(noise, axis=1) + np.finfo(float).eps)
             / 10.0 ** (snr_db / 20.0))
    return x + noise * scale[:, None], z, delta


def gen_voxels_device(n_voxels, n_scans, t_r=1.0, hrf_dur=20.0, snr_db=10.0, nb_events=5, avg_dur=12.0,
                      seed=0, first_voxel=0, delta_range=(0.7, 1.3), dtype=None, device=None,
                      return_truth=False):
    """Generate the "philox-v1" batch on the GPU (``pb_synth_voxels``): ``y[V, T]`` as a CUDA tensor,
    plus ``(z_true[V, T], delta[V])`` when ``return_truth``.  The batch that
    ``examples/icassp_2019/simulation.py:27-48`` builds voxel by voxel on the host."""
    import torch

    from . import _lib
    from ._array import stream_ptr
    dtype = dtype or torch.float32
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    V, T = int(n_voxels), int(n_scans)
    y = torch.empty((V, T), dtype=dtype, device=dev)
    z = torch.empty((V, T), dtype=dtype, device=dev) if return_truth else None
    dl = torch.empty(V, dtype=dtype, device=dev) if return_truth else None
    with torch.cuda.device(dev):
        rc = _lib.fn("pb_synth_voxels", dtype)(
            int(seed), int(first_voxel), float(t_r), float(hrf_dur), float(snr_db), int(nb_events),
            int(math.ceil(avg_dur / t_r)), float(delta_range[0]), float(delta_range[1]),
            y.data_ptr(), z.data_ptr() if return_truth else None, dl.data_ptr() if return_truth else None,
            V, T, stream_ptr())
    _lib.check(rc, "pb_synth_voxels")
    return (y, z, dl) if return_truth else y
