"""Developer check: the operator kernels (cumsum, K-tap convolution, fused, their adjoints, the layout adapter) on random
shapes against plain torch in double; exit code 1 on any mismatch."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import convolution as cv
from pybold_b200.io import voxels_from_timeseries, timeseries_from_voxels

rng = np.random.RandomState(0)
bad = 0
n = 0


def ref_conv(k, x):          # causal, truncated to T: out[t] = sum_j k[j] x[t - j]
    V, T = x.shape
    K = k.shape[-1]
    kk = k if k.dim() == 2 else k[None].expand(V, K)
    out = torch.zeros_like(x)
    for j in range(min(K, T)):
        out[:, j:] += kk[:, j:j + 1] * x[:, :T - j]
    return out


def ref_corr(k, x):          # adjoint: out[t] = sum_j k[j] x[t + j]
    V, T = x.shape
    K = k.shape[-1]
    kk = k if k.dim() == 2 else k[None].expand(V, K)
    out = torch.zeros_like(x)
    for j in range(min(K, T)):
        out[:, :T - j] += kk[:, j:j + 1] * x[:, j:]
    return out


shapes = [(1, 1, 1), (3, 2, 5), (5, 7, 3), (2, 4096, 64), (1, 4096, 1), (17, 300, 20), (4, 1200, 28), (9, 301, 20),
          (6, 296, 27), (3, 8, 8), (2, 33, 40), (5, 1000, 63)]
for _ in range(60):
    shapes.append((int(rng.randint(1, 40)), int(rng.randint(1, 4097)), int(rng.randint(1, 65))))
for V, T, K in shapes:
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 2e-5)):
        for batched in (False, True):
            x = torch.randn(V, T, device="cuda", dtype=dt)
            k = torch.randn((V, K) if batched else (K,), device="cuda", dtype=dt)
            xd, kd = x.double(), k.double()
            D = pb.DiscretInteg()
            H = pb.ConvAndLinear(D, k, dim_in=T)
            got = {"integ op": D.op(x), "integ adj": D.adj(x), "conv": cv.simple_convolve(k, x),
                   "conv adj": cv.simple_retro_convolve(k, x), "hrfinteg op": H.op(x), "hrfinteg adj": H.adj(x)}
            want = {"integ op": torch.cumsum(xd, 1), "integ adj": torch.flip(torch.cumsum(torch.flip(xd, [1]), 1), [1]),
                    "conv": ref_conv(kd, xd), "conv adj": ref_corr(kd, xd),
                    "hrfinteg op": ref_conv(kd, torch.cumsum(xd, 1)),
                    "hrfinteg adj": torch.flip(torch.cumsum(torch.flip(ref_corr(kd, xd), [1]), 1), [1])}
            for name in got:
                n += 1
                e = float((got[name].double() - want[name]).norm() / (want[name].norm() + 1e-300))
                scale = max(1.0, T ** 0.5 / 8) if dt == torch.float32 else 1.0
                if not (e < tol * scale) or tuple(got[name].shape) != (V, T):
                    bad += 1
                    print("MISMATCH %-12s V %3d T %4d K %2d %s batched=%s: %.2e" % (name, V, T, K, dt, batched, e), flush=True)
# layout adapter, both directions, odd sizes
for T, V in [(1, 1), (3, 5), (300, 65), (301, 1000), (1200, 257), (64, 64), (4096, 33), (37, 4097), (128, 2048)]:
    for dt in (torch.float32, torch.float64):
        a = torch.randn(T, V, device="cuda", dtype=dt)
        b = voxels_from_timeseries(a)
        c = timeseries_from_voxels(b)
        n += 2
        if not (torch.equal(b, a.t().contiguous()) and torch.equal(c, a)):
            bad += 1
            print("MISMATCH layout adapter T %d V %d %s" % (T, V, dt), flush=True)
print("%d checks, %d mismatches" % (n, bad))
sys.exit(1 if bad else 0)
