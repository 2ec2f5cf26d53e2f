"""Developer check: the FP64 build (tiled where a variant exists, generic elsewhere) against the CPU oracle on a spread of
series lengths and tap counts; bd with the exact theta step, deconv with early stopping.  (Oracle = test infrastructure.)"""
import sys
import numpy as np
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import _lib
from pybold_b200.hrf_model import hrf_len
from pybold_b200.synth import gen_voxels
import oracle.pybold_oracle as orc

rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (np.linalg.norm(b) + 1e-300))
bad = n = 0
worst = 0.0
for t_r in (1.0, 0.72, 0.5, 0.32):
    K = hrf_len(t_r, 20.0)
    for T in (K + 5, 97, 129, 300, 333, 640, 777, 1025, 1500, 2000):
        if T < K + 5:
            continue
        y = gen_voxels(2, T, t_r, 20.0, seed0=9300 + T)
        x, z, dz, h, d = pb.bd(y, t_r, lbda=1.1, theta_0=1.7, hrf_dur=20.0, nb_iter=3)
        xo, zo, wo, ho, do = orc.bd(y[1], t_r, lbda=1.1, theta_0=1.7, hrf_dur=20.0, nb_iter=3, theta_solver="exact")
        e = max(rel(z[1], zo), rel(h[1], ho), rel(d["J"][1], do["J"]), rel(x[1], xo))
        worst = max(worst, e)
        n += 1
        if not e < 2e-8:
            bad += 1
            print("MISMATCH bd T %d K %d variant %d: %.2e" % (T, K, _lib.lib.pb_solver_variant(T, K, 1), e), flush=True)
        hh, _ = orc.spm_hrf(1.0, t_r, 20.0, True)
        x0 = np.random.RandomState(T).randn(T)
        r = pb.deconv(y[0], t_r, hh, lbda=0.9, nb_iter=60, early_stopping=True, tol=5e-2, x0=x0)
        Lc = 0.9 * orc.spectral_radius_est(orc.HrfIntegOperator(hh, T), x0)
        xo, zo, wo, Jo, n_o = orc.deconv_fixed_lbda(y[0], hh, 0.9, lipschitz=Lc, early_stopping=True, tol=5e-2, wind=6, nb_iter=60)
        e = max(rel(r[1], zo), rel(r[3], Jo) if len(r[3]) == n_o else 1.0)
        worst = max(worst, e)
        n += 1
        if not e < 1e-9:
            bad += 1
            print("MISMATCH deconv T %d K %d: %.2e (iterations %d vs %d)" % (T, K, e, len(r[3]), n_o), flush=True)
print("%d checks, %d mismatches, worst relative error %.1e" % (n, bad, worst))
sys.exit(1 if bad else 0)
