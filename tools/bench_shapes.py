"""Developer measurement: bd throughput over series lengths and tap counts (which register-tiled variant
the dispatcher picks and what fraction of the nominal FP32 peak it delivers).

    python tools/bench_shapes.py [nb_iter] [waves] > profiles/rNN_shapes.txt

Every shape runs whole waves of its persistent grid (no tail effect); flops as in bench.py.
"""
import sys
import torch
sys.path.insert(0, ".")
from bench import flops_bd_voxel
from pybold_b200 import _lib
from pybold_b200.bold_signal import bd_alloc, bd_batch
from pybold_b200.hrf_model import hrf_len
from pybold_b200.synth import gen_voxels_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sms = torch.cuda.get_device_properties(0).multi_processor_count
nominal = sms * 128 * 2 * 1.965e9 / 1e12
Ts = [64, 96, 100, 128, 150, 190, 200, 240, 256, 300, 320, 350, 384, 405, 450, 500, 512, 600, 650, 700, 768, 800,
      900, 1000, 1050, 1150, 1200, 1280, 1500, 2000, 2400, 2560, 3000, 4096]
import os
if os.environ.get("PB_SHAPES_TS"):          # e.g. PB_SHAPES_TS=2600,3000,4096 PB_SHAPES_TR=1.0,0.72
    Ts = [int(v) for v in os.environ["PB_SHAPES_TS"].split(",")]
TRs = [float(v) for v in os.environ.get("PB_SHAPES_TR", "1.0,0.75,0.72,0.5,0.32").split(",")]
print("bd, FP32, nb_iter = %d, %d waves of the grid per shape; fraction of the nominal FP32 peak (%.1f Tflop/s)" % (n, waves, nominal))
print("%5s %3s %10s %8s %9s %11s %8s %6s" % ("T", "K", "variant", "voxels", "ms", "voxels/s", "Tflop/s", "frac"))
worst = {}
for t_r in TRs:
    K = hrf_len(t_r, 20.0)
    for T in Ts:
        wave = _lib.lib.pb_bd_wave_voxels(T, K, 0, n)
        V = waves * wave if wave > 0 else 2000
        y = gen_voxels_device(V, T, t_r, 20.0)
        out = bd_alloc(V, T, K, n, torch.float32, y.device)
        lb = torch.full((1,), 1.7, device="cuda"); th = torch.full((1,), 2.0, device="cuda")
        f = lambda: bd_batch(y, t_r, lb, th, None, 20.0, [(0.6, 1.9)], n, False, 4, 1e-12, out=out)
        f(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf = flops_bd_voxel(T, K, n) * V / best / 1e9
        vid = _lib.lib.pb_solver_variant(T, K, 0)
        print("%5d %3d %10d %8d %9.2f %11.0f %8.1f %6.2f" % (T, K, vid, V, best, V / best * 1e3, tf, tf / nominal), flush=True)
        if T <= 2560:
            worst[K] = min(worst.get(K, 9), tf / nominal)
        del y, out
print("lowest fraction for 64 <= T <= 2560 per tap count:", {k: round(v, 2) for k, v in worst.items()})
