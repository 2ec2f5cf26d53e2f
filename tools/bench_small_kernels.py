"""Developer measurement: the one-warp-per-voxel helper kernels (hrf_estim, power iteration, Frobenius Lipschitz, sigma, spm_hrf) at
the cfg3 batch size, against the bytes they have to read (context for DESIGN.md section 9: how far from the solve they are)."""
import sys
import torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.bold_signal import hrf_estim_batch, bd_batch
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
y, z, _ = gen_voxels_device(V, T, t_r, 20.0, return_truth=True)
ms = timed(lambda: hrf_estim_batch(z, y, t_r, 20.0))
print("hrf_estim      %d x %d: %8.3f ms  (reads z, y: %.0f MB -> %.3f ms at 6554 GB/s)" % (V, T, ms, 2 * V * T * 4 / 1e6, 2 * V * T * 4 / 6554e6))
h = torch.as_tensor(pb.spm_hrf(1.0, t_r, 20.0)[0], dtype=torch.float32, device="cuda")
hb = h[None].repeat(V, 1).contiguous()
x0 = torch.randn(V, T, device="cuda")
ms = timed(lambda: pb.utils.spectral_radius_est(pb.ConvAndLinear(pb.DiscretInteg(), hb, dim_in=T), (T,), x0=x0))
print("power iteration %d x %d (30 steps, per-voxel taps): %8.3f ms" % (V, T, ms))
ms = timed(lambda: pb.utils.mad_daub_noise_est(y))
print("mad_daub_noise_est %d x %d: %8.3f ms  (reads y: %.3f ms at 6554 GB/s)" % (V, T, ms, V * T * 4 / 6554e6))
th = torch.rand(V, device="cuda") * 1.2 + 0.6
ms = timed(lambda: pb.spm_hrf(th, t_r, 20.0))
print("spm_hrf %d thetas: %8.3f ms" % (V, ms))
lb = torch.full((1,), 1.7, device="cuda"); t0 = torch.full((1,), 2.0, device="cuda")
ms = timed(lambda: bd_batch(y, t_r, lb, t0, None, 20.0, [(0.6, 1.9)], 100, False, 4, 1e-12), reps=1)
print("bd (nb_iter = 100) for scale: %8.1f ms" % ms)
