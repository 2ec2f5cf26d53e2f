"""Runs the UNMODIFIED reference (hcherkaoui/pybold) for ``bench.py``'s CPU arms.

``baseline/_ref/`` (git-ignored, shipped to the GPU box with the working tree) holds the reference as
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``
installed it (``__graft_entry__.build()`` repeats the install when ``/root/reference`` is present and
``baseline/_ref`` is not).  Nothing of it is modified; three things are arranged AROUND it, all needed
just to import it with this image's NumPy 2.3 and without PyWavelets:

 * ``pywt`` stub module -- ``pybold/utils.py:7`` imports PyWavelets at module top; only
   ``deconv(lbda=None)`` calls it and the benchmark never does;
 * ``np.float = float`` -- removed in NumPy 1.24, used by ``pybold/data.py:441`` (import-time safe,
   kept for completeness);
 * ``NUMBA_CACHE_DIR`` -> a writable directory (``_loops_deconv`` is ``@jit(cache=True)``).

This module imports neither ``pybold_b200`` nor ``oracle``: the reference arm runs none of our code
inside its timed region.  The fan-out is the reference's own pattern
(``joblib.Parallel`` over voxels, examples/icassp_2019/validation.py:43-47) with BLAS / OpenMP pinned to
one thread per worker as its examples do (validation.py:6-9).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_bs = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, "pybold", "bold_signal.py"))


def pin_threads():
    """One BLAS / OpenMP thread per worker process; call BEFORE NumPy is imported in the parent so
    that the joblib workers inherit it."""
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[k] = "1"
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/pybold_ref_numba_cache")
    root = os.path.dirname(HERE)
    pp = os.environ.get("PYTHONPATH", "")
    if root not in pp.split(os.pathsep):
        os.environ["PYTHONPATH"] = root + (os.pathsep + pp if pp else "")


def load():
    """Import ``pybold.bold_signal`` of the reference (once per process)."""
    global _bs
    if _bs is not None:
        return _bs
    if not available():
        raise ImportError("baseline/_ref/pybold is missing (see the module docstring)")
    import numpy as np
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/pybold_ref_numba_cache")
    os.makedirs(os.environ["NUMBA_CACHE_DIR"], exist_ok=True)
    if not hasattr(np, "float"):
        np.float = float
    sys.modules.setdefault("pywt", types.ModuleType("pywt"))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import pybold.bold_signal as bs
    _bs = bs
    return bs


def bd_one(yv, t_r, lbda, theta_0, hrf_dur, bounds, nb_iter):
    """``pybold.bold_signal.bd`` on one voxel; returns the final theta (so the call cannot be elided)."""
    bs = load()
    with contextlib.redirect_stdout(io.StringIO()):
        x, z, dz, h, d = bs.bd(yv, t_r=t_r, lbda=lbda, theta_0=theta_0, hrf_dur=hrf_dur,
                               bounds=[tuple(bounds)], nb_iter=nb_iter)
    return float(d["J"][-1])


def deconv_one(yv, t_r, hrf, lbda, nb_iter):
    """``pybold.bold_signal.deconv`` (fixed lambda, no early stopping) on one voxel; the reference
    prints every iteration (bold_signal.py:79-80), hence the redirect."""
    bs = load()
    with contextlib.redirect_stdout(io.StringIO()):
        out = bs.deconv(yv, t_r=t_r, hrf=hrf, lbda=lbda, early_stopping=False, nb_iter=nb_iter)
    return float(out[3][-1])


def bd_rate(y_sample, w, n_jobs, pool=None):
    """(voxels/s, seconds) of the reference's ``bd`` over ``y_sample`` with ``n_jobs`` workers."""
    from joblib import Parallel, delayed
    par = pool if pool is not None else Parallel(n_jobs=n_jobs)
    t0 = time.perf_counter()
    par(delayed(bd_one)(yv, w["t_r"], w["lbda"], w["theta_0"], w["hrf_dur"], w["bounds"], w["nb_iter"])
        for yv in y_sample)
    dt = time.perf_counter() - t0
    return len(y_sample) / dt, dt
