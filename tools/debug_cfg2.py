import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.bold_signal import deconv_batch
from pybold_b200.synth import gen_voxels_device
dev = torch.device("cuda", 0)
y2 = gen_voxels_device(10000, 300, 1.0, 20.0, seed=2, dtype=torch.float32)
h2 = torch.as_tensor(pb.spm_hrf(1.0, 1.0, 20.0, True)[0], dtype=torch.float32, device=dev)
L2t = torch.full((1,), 3.0e5, device=dev); lb2 = torch.full((1,), 1.0, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for trial in range(3):
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        out = deconv_batch(y2, h2, lb2, L2t, None, False, 1e-6, 6, 200)
        t1 = time.perf_counter()
        flush.fill_(1.0)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print("host ms per call (issue, issue+sync):", [("%.2f" % a, "%.2f" % b) for a, b in ts])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    out = deconv_batch(y2, h2, lb2, L2t, None, False, 1e-6, 6, 200)
torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
