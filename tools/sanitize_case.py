"""Tiny run of every solver kernel family (for compute-sanitizer memcheck / racecheck)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.synth import gen_voxels
from pybold_b200 import convolution as cv
from pybold_b200.synth import gen_voxels_device
from pybold_b200.utils import inf_norm, rel_l2_err
from pybold_b200.io import voxels_from_timeseries
for (T, t_r, dt) in [(300, 1.0, np.float32), (240, 0.75, np.float32), (600, 1.0, np.float32),
                     (1200, 0.72, np.float32), (100, 1.0, np.float32), (200, 0.5, np.float32),
                     (150, 1.0, np.float32), (350, 1.0, np.float32), (700, 0.72, np.float32),
                     (2000, 0.72, np.float32), (300, 1.0, np.float64), (1200, 0.72, np.float64),
                     # round 2: six-warp CTAs, gap-filling 40- / 64-tap variants, short series with 28 taps
                     (3000, 1.0, np.float32), (3500, 0.72, np.float32), (3600, 0.5, np.float32), (4096, 1.0, np.float32),
                     (450, 0.5, np.float32), (700, 0.5, np.float32), (900, 0.5, np.float32), (1500, 0.5, np.float32),
                     (1800, 0.5, np.float32), (200, 0.32, np.float32), (450, 0.32, np.float32), (700, 0.32, np.float32),
                     (900, 0.32, np.float32), (1500, 0.32, np.float32), (96, 0.72, np.float32), (128, 0.72, np.float32),
                     (450, 1.0, np.float32), (1000, 0.72, np.float32), (1500, 1.0, np.float32)]:
    y = gen_voxels(5, T, t_r, 20.0, seed0=1).astype(dt)
    out = pb.bd(y, t_r, lbda=1.0, nb_iter=3)
    out = pb.bd(y, t_r, lbda=1.0, nb_iter=6, early_stopping=True, tol=1e-2)
    h, _ = pb.spm_hrf(1.0, t_r, 20.0)
    out = pb.deconv(y, t_r, h.astype(dt), lbda=0.5, nb_iter=12, early_stopping=True, tol=1e-3, x0=np.ones(T, dtype=dt))
    lp = pb.bold_signal.deconv_lbda_path(y, t_r, h.astype(dt), [0.3, 0.6], nb_iter=5) if T <= 1200 else None
    au = pb.deconv(y[:2], t_r, h.astype(dt), lbda=None, sigma=0.5, nb_iter=3, nb_sub_iter=5) if T <= 600 else None
    print("ok", T, t_r, dt.__name__, float(np.abs(out[1]).max()), flush=True)
# operator kernels (register-resident rows and the shared-memory fallback), N3, N4
for (V, T, K) in [(7, 300, 20), (5, 1200, 28), (3, 301, 20), (4, 600, 40), (9, 4, 1)]:
    x = torch.randn(V, T, device="cuda")
    k = torch.randn(V, K, device="cuda")
    D = pb.DiscretInteg()
    H = pb.ConvAndLinear(D, k, dim_in=T)
    for r in (D.op(x), D.adj(x), cv.simple_convolve(k, x), cv.simple_retro_convolve(k, x), H.op(x), H.adj(x)):
        assert r.shape == (V, T)
    print("ops ok", V, T, K)
y, z, dl = gen_voxels_device(33, 300, return_truth=True)
e = rel_l2_err(inf_norm(y), inf_norm(z))
vt = voxels_from_timeseries(torch.randn(300, 65, device="cuda"))
torch.cuda.synchronize()
print("done", float(e.mean()), tuple(vt.shape))
