"""Tiny run of every solver kernel family (for compute-sanitizer memcheck / racecheck)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.synth import gen_voxels
for (T, t_r, dt) in [(300, 1.0, np.float32), (240, 0.75, np.float32), (600, 1.0, np.float32),
                     (1200, 0.72, np.float32), (100, 1.0, np.float32), (200, 0.5, np.float32),
                     (300, 1.0, np.float64), (1200, 0.72, np.float64)]:
    y = gen_voxels(5, T, t_r, 20.0, seed0=1).astype(dt)
    out = pb.bd(y, t_r, lbda=1.0, nb_iter=3)
    out = pb.bd(y, t_r, lbda=1.0, nb_iter=6, early_stopping=True, tol=1e-2)
    h, _ = pb.spm_hrf(1.0, t_r, 20.0)
    out = pb.deconv(y, t_r, h.astype(dt), lbda=0.5, nb_iter=12, early_stopping=True, tol=1e-3, x0=np.ones(T, dtype=dt))
    print("ok", T, t_r, dt.__name__, float(np.abs(out[1]).max()))
torch.cuda.synchronize()
print("done")
