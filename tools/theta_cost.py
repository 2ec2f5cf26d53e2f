"""Developer tool: how much of the bd kernel time is the theta phase? (bounds pinned vs free)"""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from pybold_b200.bold_signal import bd_alloc, bd_batch
from pybold_b200.synth import gen_voxels_chunked
V, T = 42624, 300
y = torch.as_tensor(gen_voxels_chunked(V, T), device="cuda")
out = bd_alloc(V, T, 20, 100, torch.float32, y.device)
lb = torch.full((1,), 1.7, device="cuda"); th = torch.full((1,), 2.0, device="cuda")
def t(bounds, label):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); bd_batch(y, 1.0, lb, th, None, 20.0, bounds, 100, False, 4, 1e-12, out=out); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("%-40s %.1f ms  %.0f voxels/s" % (label, best, V / best * 1e3))
t([(0.6, 1.9)], "free theta")
t([(1.0, 1.0)], "theta pinned (1 eval per outer iteration)")
