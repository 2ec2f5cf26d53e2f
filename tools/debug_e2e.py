"""Developer probe: per-call wall time of the e2e leg (pb.bd on a pinned host batch), with CUDA events around
every chunk's solver launch (monkeypatched bd_batch) -- where does a slow call lose its time?"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import bold_signal as bs
from pybold_b200.synth import gen_voxels_chunked

V, T = 100000, 300
y_host = torch.from_numpy(gen_voxels_chunked(V, T, 1.0, 20.0, dtype=np.float32)).pin_memory()
marks = []
orig = bs.bd_batch


def traced(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host = time.perf_counter()
    e0.record()
    r = orig(*a, **k)
    e1.record()
    marks.append((e0, e1, t_host, time.perf_counter()))
    return r


bs.bd_batch = traced
fine = []
orig_pv = bs.per_voxel


def pv(*a, **k):
    t = time.perf_counter()
    r = orig_pv(*a, **k)
    fine.append(("per_voxel", (time.perf_counter() - t) * 1e3))
    return r


bs.per_voxel = pv
from pybold_b200 import _lib
orig_fn = _lib.fn


def fn(name, dtype):
    f0 = orig_fn(name, dtype)

    def call(*a):
        t = time.perf_counter()
        r = f0(*a)
        fine.append((name, (time.perf_counter() - t) * 1e3))
        return r
    return call


_lib.fn = fn
f = lambda: pb.bd(y_host, 1.0, lbda=1.7, theta_0=2.0, hrf_dur=20.0, bounds=[(0.6, 1.9)], nb_iter=100)
res = None
for i in range(14):
    del marks[:]
    del fine[:]
    g0 = torch.cuda.Event(enable_timing=True); g0.record()
    t0 = time.perf_counter()
    res = f()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ks = ["%.1f" % a.elapsed_time(b) for a, b, _, _ in marks]
    starts = ["%.1f" % g0.elapsed_time(a) for a, _, _, _ in marks]
    ends = ["%.1f" % g0.elapsed_time(b) for _, b, _, _ in marks]
    host = ["%.1f..%.1f" % ((a - t0) * 1e3, (b - t0) * 1e3) for _, _, a, b in marks]
    print("call %2d: %7.1f ms | kernel ms %s | start %s | end %s | host launch %s" % (i, (t1 - t0) * 1e3, ks, starts, ends, host), flush=True)
    if (t1 - t0) > 0.62:
        print("   fine:", ["%s %.2f" % v for v in fine], flush=True)
