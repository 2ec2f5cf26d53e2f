"""Developer measurement: bd throughput at series lengths between the tuned shapes (which register-tiled
variant the dispatcher picks and what it delivers)."""
import sys, torch, json
sys.path.insert(0, ".")
from pybold_b200 import _lib
from pybold_b200.bold_signal import bd_alloc, bd_batch
from pybold_b200.hrf_model import hrf_len
from pybold_b200.synth import gen_voxels_device
for T, t_r, V in [(100, 1.0, 80000), (128, 0.72, 60000), (150, 1.0, 60000), (190, 1.0, 50000), (350, 1.0, 40000), (330, 0.72, 40000), (405, 1.0, 30000), (500, 0.72, 24000), (650, 1.0, 20000),
                  (700, 0.72, 16000), (800, 0.72, 16000), (900, 0.72, 14000), (1000, 0.72, 12000), (1050, 0.72, 12000), (1150, 0.72, 12000), (2000, 0.72, 6000), (2400, 1.0, 6000)]:
    K = hrf_len(t_r, 20.0)
    y = gen_voxels_device(V, T, t_r, 20.0)
    out = bd_alloc(V, T, K, 100, torch.float32, y.device)
    lb = torch.full((1,), 1.7, device="cuda"); th = torch.full((1,), 2.0, device="cuda")
    f = lambda: bd_batch(y, t_r, lb, th, None, 20.0, [(0.6, 1.9)], 100, False, 4, 1e-12, out=out)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    mac = T * K - K * (K - 1) // 2
    fl = V * (101 * 100 * (4 * mac + 11 * T))
    print(T, K, "variant", _lib.lib.pb_solver_variant(T, K, 0), "%.1f ms %.0f vox/s %.1f Tflop/s" % (ms, V / ms * 1e3, fl / ms / 1e9), flush=True)
