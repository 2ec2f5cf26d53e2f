"""Developer check: for every FP32 variant, a voxel's bd result inside a batch (several voxels per warp, work queue) is bit-identical
to the same voxel solved alone, with per-voxel lbda / theta_0 and a warm start; the same for deconv with early stopping."""
import sys
import numpy as np
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import _lib
from pybold_b200.hrf_model import hrf_len
from pybold_b200.synth import gen_voxels

bad = n = 0
for t_r in (1.0, 0.72, 0.5, 0.32):
    K = hrf_len(t_r, 20.0)
    seen = {}
    for T in range(K + 4, 4097):
        vid = _lib.lib.pb_solver_variant(T, K, 0)
        if vid and vid not in seen:
            seen[vid] = T
    h = pb.spm_hrf(1.0, t_r, 20.0)[0].astype(np.float32)
    for vid, T0 in sorted(seen.items()):
        T = min(4096, T0 + 3)
        if _lib.lib.pb_solver_variant(T, K, 0) != vid:
            T = T0
        V = 37
        rs = np.random.RandomState(T)
        y = (gen_voxels(V, T, t_r, 20.0, seed0=9900 + T) * rs.uniform(0.5, 2.0, (V, 1))).astype(np.float32)
        lb = rs.uniform(0.5, 2.0, V)
        th = rs.uniform(0.7, 1.8, V)
        z0 = np.zeros((V, T), dtype=np.float32)
        z0[:, T // 3:T // 2] = 1.0
        for es in (False, True):
            kw = dict(hrf_dur=20.0, nb_iter=4, early_stopping=es, tol=1e-2)
            x, z, dz, hh, d = pb.bd(y, t_r, lbda=lb, theta_0=th, z_0=z0, **kw)
            for v in (0, 17, V - 1):
                x1, z1, dz1, h1, d1 = pb.bd(y[v], t_r, lbda=float(lb[v]), theta_0=float(th[v]), z_0=z0[v], **kw)
                n += 1
                nt = len(d1["J"])
                ok = np.array_equal(z1, z[v]) and np.array_equal(h1, hh[v]) and np.array_equal(x1, x[v]) and \
                    np.array_equal(np.asarray(d1["J"]), np.asarray(d["J"][v])[:nt])
                if not ok:
                    bad += 1
                    print("MISMATCH bd variant %d T %d K %d es %s voxel %d" % (vid, T, K, es, v), flush=True)
        x, z, dz, J, _, _ = pb.deconv(y, t_r, h, lbda=0.9, nb_iter=40, early_stopping=True, tol=2e-2, x0=np.ones(T, dtype=np.float32))
        for v in (0, 17, V - 1):
            x1, z1, dz1, J1, _, _ = pb.deconv(y[v], t_r, h, lbda=0.9, nb_iter=40, early_stopping=True, tol=2e-2,
                                              x0=np.ones(T, dtype=np.float32))
            n += 1
            if not (np.array_equal(z1, z[v]) and len(J1) == int(np.sum(~np.isnan(J[v])))):
                bad += 1
                print("MISMATCH deconv T %d K %d voxel %d" % (T, K, v), flush=True)
print("%d checks, %d mismatches" % (n, bad))
sys.exit(1 if bad else 0)
