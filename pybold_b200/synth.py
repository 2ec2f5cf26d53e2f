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
