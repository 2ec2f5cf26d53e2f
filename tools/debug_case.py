"""Developer check: FP32 variant vs FP64 generic kernel on long series; theta free / theta fixed."""
import sys
import numpy as np
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.synth import gen_voxels

def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)

for T, t_r, n_it in [(2000, 0.72, 4), (4096, 0.72, 3), (4096, 0.72, 1), (1200, 0.72, 10)]:
    y = gen_voxels(3, T, t_r, 20.0, seed0=7500 + T)
    for bounds, th0, tag in (([(0.6, 1.9)], 2.0, "theta free "), ([(1.0, 1.0)], 1.0, "theta fixed")):
        x, z, dz, h, d = pb.bd(y, t_r, lbda=1.4, theta_0=th0, hrf_dur=20.0, nb_iter=n_it, bounds=bounds)
        x32, z32, dz32, h32, d32 = pb.bd(y.astype(np.float32), t_r, lbda=1.4, theta_0=th0, hrf_dur=20.0, nb_iter=n_it, bounds=bounds)
        print(T, n_it, tag, "z %.2e dz %.2e x %.2e h %.2e J %.2e theta %.2e" % (rel(z32, z), rel(dz32, dz), rel(x32, x), rel(h32, h), rel(d32["J"], d["J"]),
              np.max(np.abs(d32["theta"] - d["theta"]))), flush=True)
