"""Developer measurement (round 2): work queue on / off for the bd kernel, cfg2 deconv launch timing."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.bold_signal import bd_alloc, bd_batch, deconv_batch
from pybold_b200.hrf_model import hrf_len, spm_hrf
from pybold_b200.synth import gen_voxels_device


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


V, T, t_r = 100000, 300, 1.0
K = hrf_len(t_r, 20.0)
y = gen_voxels_device(V, T, t_r, 20.0)
out = bd_alloc(V, T, K, 100, torch.float32, y.device)
lb = torch.full((1,), 1.7, device="cuda"); th = torch.full((1,), 2.0, device="cuda")
for mode in ("queue", "static", "queue", "static"):
    if mode == "static":
        os.environ["PB_NO_QUEUE"] = "1"
    else:
        os.environ.pop("PB_NO_QUEUE", None)
    ms = timed(lambda: bd_batch(y, t_r, lb, th, None, 20.0, [(0.6, 1.9)], 100, False, 4, 1e-12, out=out))
    print("bd 100k x 300 (%s): %.2f ms  %.1f k voxels/s" % (mode, ms, V / ms))
os.environ.pop("PB_NO_QUEUE", None)

h = torch.as_tensor(spm_hrf(1.0, t_r, 20.0, True)[0], device="cuda", dtype=torch.float32)
for Vd in (10000, 14208, 50000, 100000):
    yd = gen_voxels_device(Vd, T, t_r, 20.0, seed=2)
    lbt = torch.full((1,), 1.0, device="cuda"); Lt = torch.full((1,), 3.0e5, device="cuda")
    ms = timed(lambda: deconv_batch(yd, h, lbt, Lt, None, False, 1e-6, 6, 200), reps=5)
    print("deconv %d x 300, 200 it (tensor params): %.3f ms  %.2f M voxels/s" % (Vd, ms, Vd / ms / 1e3))
    ms = timed(lambda: deconv_batch(yd, h, 1.0, 3.0e5, None, False, 1e-6, 6, 200), reps=5)
    print("deconv %d x 300, 200 it (scalar params): %.3f ms  %.2f M voxels/s" % (Vd, ms, Vd / ms / 1e3))
