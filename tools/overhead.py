import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import _lib
from pybold_b200.bold_signal import bd_batch
from pybold_b200.synth import gen_voxels_chunked
T=300
yall = torch.as_tensor(gen_voxels_chunked(40000, T), device="cuda")
for V in [1776, 3552, 7104, 20000, 40000, 1776, 20000]:
    y = yall[:V].contiguous()
    for rep in range(3):
        torch.cuda.synchronize()
        t0=time.perf_counter()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        out = bd_batch(y, 1.0, 1.7, 2.0, None, 20.0, [(0.6, 1.9)], 100, False, 4, 1e-12)
        t1=time.perf_counter()
        e1.record(); torch.cuda.synchronize()
        t2=time.perf_counter()
        print("V=%6d rep%d events %.2f ms  host-enqueue %.2f ms  wall %.2f ms" % (V, rep, e0.elapsed_time(e1), (t1-t0)*1e3, (t2-t0)*1e3))
