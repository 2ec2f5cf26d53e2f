"""Developer measurement: end-to-end bd() on a pinned host batch (the public call of bench.py's e2e leg)
against the device-resident launch, chunk sizes of the streamed path."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import bold_signal as bs
from pybold_b200.bold_signal import bd_alloc, bd_batch
from pybold_b200.synth import gen_voxels_chunked

V, T = 100000, 300
y_host = torch.from_numpy(gen_voxels_chunked(V, T, 1.0, 20.0, dtype=np.float32)).pin_memory()
y_dev = y_host.cuda()
out = bd_alloc(V, T, 20, 100, torch.float32, y_dev.device)
lb = torch.full((1,), 1.7, device="cuda"); th = torch.full((1,), 2.0, device="cuda")


def dev_ms():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bd_batch(y_dev, 1.0, lb, th, None, 20.0, [(0.6, 1.9)], 100, False, 4, 1e-12, out=out); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def e2e_ms(n=4):
    f = lambda: pb.bd(y_host, 1.0, lbda=1.7, theta_0=2.0, hrf_dur=20.0, bounds=[(0.6, 1.9)], nb_iter=100)
    r = f(); r = f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


dev_ms()
print("device-resident launch: %.1f ms" % min(dev_ms(), dev_ms()))
for chunk in (32768, 16384, 65536, 8192, 100000):
    bs._STREAM_TARGET_CHUNK = chunk
    print("e2e, chunk target %6d: %.1f ms" % (chunk, e2e_ms()), flush=True)
bs._STREAM_TARGET_CHUNK = 32768
os.environ["PB_NO_QUEUE"] = "1"
print("static scheduling: device %.1f ms, e2e %.1f ms" % (min(dev_ms(), dev_ms()), e2e_ms()))
