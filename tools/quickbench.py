"""Quick device timing of bd / deconv batches (developer tool, not the bench contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import _lib
from pybold_b200.bold_signal import bd_batch, deconv_batch
from pybold_b200.synth import gen_voxels

def run(V, T, t_r, nb_iter, dtype=torch.float32, reps=2):
    y0 = gen_voxels(min(V, 2048), T, t_r, 20.0, seed0=0)
    y = torch.as_tensor(np.tile(y0, (V // len(y0) + 1, 1))[:V], device="cuda", dtype=dtype).contiguous()
    K = _lib.lib.pb_hrf_len(t_r, 20.0)
    var = _lib.lib.pb_solver_variant(T, K, int(dtype == torch.float64))
    for _ in range(1):
        bd_batch(y[:4096], t_r, 1.7, 2.0, None, 20.0, [(0.6, 1.9)], nb_iter, False, 4, 1e-12)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        out = bd_batch(y, t_r, 1.7, 2.0, None, 20.0, [(0.6, 1.9)], nb_iter, False, 4, 1e-12)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    MAC = T * K - K * (K - 1) // 2
    F_it = 4 * MAC + 11 * T
    flops = V * (nb_iter + 1) * nb_iter * F_it
    print("bd V=%d T=%d K=%d nb_iter=%d %s variant=%d: %.1f ms  %.0f voxels/s  %.2f Tflop/s (inner-loop algorithmic)"
          % (V, T, K, nb_iter, str(dtype).split('.')[-1], var, best, V / best * 1e3, flops / best / 1e9))

if __name__ == "__main__":
    run(20000, 300, 1.0, 100)
    run(4000, 1200, 0.72, 100)
    run(20000, 240, 0.75, 100)
    run(4000, 300, 1.0, 100, torch.float64)
