"""Developer measurement: cost of the early-stopping kernels (one warp per voxel) next to the group
kernels that serve the same shapes without early stopping (tol = 0: nothing stops early)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pybold_b200.bold_signal import bd_alloc, bd_batch, deconv_batch
from pybold_b200.hrf_model import hrf_len, spm_hrf
from pybold_b200.synth import gen_voxels_device


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


V, T, t_r = 50000, 300, 1.0
K = hrf_len(t_r, 20.0)
y = gen_voxels_device(V, T, t_r, 20.0)
h = torch.as_tensor(spm_hrf(1.0, t_r, 20.0, True)[0], device="cuda", dtype=torch.float32)
for es in (False, True):
    ms = timed(lambda: deconv_batch(y, h, 1.0, 3.0e5, None, es, 0.0, 6, 200))
    print("deconv 50k x 300, 200 it, early_stopping=%s: %.2f ms  %.2f M voxels/s" % (es, ms, V / ms / 1e3))
out = bd_alloc(V, T, K, 100, torch.float32, y.device)
lb = torch.full((1,), 1.7, device="cuda"); th = torch.full((1,), 2.0, device="cuda")
for es in (False, True):
    ms = timed(lambda: bd_batch(y, t_r, lb, th, None, 20.0, [(0.6, 1.9)], 100, es, 4, 0.0, out=out))
    print("bd 50k x 300, nb_iter 100, early_stopping=%s: %.1f ms  %.1f k voxels/s" % (es, ms, V / ms))
