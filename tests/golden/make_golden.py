"""Generate the golden vectors in this directory from the LIVE reference.

Run in the build container only (needs the read-only checkout at /root/reference):

    python tests/golden/make_golden.py

The reference is imported unmodified with two shims the survey found necessary
(a stub ``pywt`` module because ``pybold/utils.py:7`` imports it at module top, and
``np.float`` which NumPy >= 1.24 removed).  Nothing from the reference is copied: only
its numerical outputs on seeded inputs are stored.  Versions at generation time are
written into each file (numpy / scipy / numba) because the L-BFGS-B theta step is
third-party SciPy code.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
if not hasattr(np, "float"):
    np.float = float
sys.modules.setdefault("pywt", types.ModuleType("pywt"))
sys.path.insert(0, "/root/reference")

import numba  # noqa: E402
import scipy  # noqa: E402
import pybold.bold_signal as ref_bs  # noqa: E402
import pybold.convolution as ref_cv  # noqa: E402
import pybold.hrf_model as ref_hm  # noqa: E402
import pybold.linear as ref_lin  # noqa: E402
import pybold.utils as ref_ut  # noqa: E402
from pybold_b200.synth import gen_voxels  # noqa: E402

VERSIONS = np.array([np.__version__, scipy.__version__, numba.__version__])


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):   # deconv prints every iteration (Q4)
        return fn(*a, **k)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, versions=VERSIONS, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def gen_ops():
    rng = np.random.RandomState(1)
    out = {}
    for idx, (T, K) in enumerate([(64, 5), (300, 20), (257, 27), (40, 40)]):
        k = rng.randn(K)
        x = rng.randn(T)
        out["k%d" % idx] = k
        out["x%d" % idx] = x
        out["conv%d" % idx] = ref_cv.simple_convolve(k, x)
        out["corr%d" % idx] = ref_cv.simple_retro_convolve(k, x)
        out["sconv%d" % idx] = ref_cv.spectral_convolve(k, x)
        out["scorr%d" % idx] = ref_cv.spectral_retro_convolve(k, x)
        Kmat = ref_cv.toeplitz_from_kernel(k, T, T)
        out["toep_dot%d" % idx] = Kmat.dot(x)
        out["toepT_dot%d" % idx] = Kmat.T.dot(x)
        H = ref_lin.ConvAndLinear(ref_lin.DiscretInteg(), k, dim_in=T, dim_out=T)
        out["H_op%d" % idx] = H.op(x)
        out["H_adj%d" % idx] = H.adj(x)
        out["integ_op%d" % idx] = ref_lin.DiscretInteg().op(x)
        out["integ_adj%d" % idx] = ref_lin.DiscretInteg().adj(x)
    save("ops", n_cases=np.array(4), **out)


HRF_GRID = [(1.0, 1.0, 20.0), (2.0, 1.0, 20.0), (0.7, 0.75, 20.0), (1.9, 0.72, 20.0),
            (0.6, 0.7535, 20.0), (1.5, 1.0, 30.0), (0.5, 2.0, 60.0), (1.3, 0.1, 10.0)]


def gen_hrf():
    out = {"grid": np.array(HRF_GRID)}
    for idx, (delta, t_r, dur) in enumerate(HRF_GRID):
        h, t = ref_hm.spm_hrf(delta, t_r, dur, False)
        hn, _ = ref_hm.spm_hrf(delta, t_r, dur, True)
        out["h%d" % idx] = h
        out["hn%d" % idx] = hn
        out["t%d" % idx] = t
    save("spm_hrf", **out)


def gen_lipschitz():
    out = {}
    cases = [(300, 1.0, 1.0, 20.0), (240, 0.7, 0.75, 20.0), (600, 1.5, 1.0, 30.0), (50, 1.0, 1.0, 20.0)]
    out["cases"] = np.array(cases)
    for idx, (T, delta, t_r, dur) in enumerate(cases):
        T = int(T)
        h, _ = ref_hm.spm_hrf(delta, t_r, dur, True)
        H = ref_lin.ConvAndLinear(ref_lin.DiscretInteg(), h, dim_in=T, dim_out=T)
        np.random.seed(100 + idx)
        x0 = np.random.randn(T)
        np.random.seed(100 + idx)
        out["x0_%d" % idx] = x0
        out["h%d" % idx] = h
        out["power%d" % idx] = np.array(ref_ut.spectral_radius_est(H, (T,)))
        # Frobenius norm of the Gram matrix exactly as _loops_deconv forms it (bold_signal.py:249-253)
        Kmat = ref_cv.toeplitz_from_kernel(h, T, T)
        A = Kmat.dot(np.tril(np.ones((T, T)), 0))
        out["frob%d" % idx] = np.array(np.linalg.norm(A.T.dot(A)))
    save("lipschitz", **out)


def gen_deconv():
    T, t_r, dur = 300, 1.0, 20.0
    y = gen_voxels(3, T, t_r, dur, seed0=7)
    h, _ = ref_hm.spm_hrf(1.0, t_r, dur, True)
    out = {"y": y, "h": h}
    runs = [("a", 0, dict(lbda=1.0, nb_iter=200, early_stopping=False)),
            ("b", 1, dict(lbda=0.3, nb_iter=120, early_stopping=False)),
            ("c", 2, dict(lbda=1.0, nb_iter=400, early_stopping=True, tol=1.0e-2)),
            ("d", 0, dict(lbda=1.0, nb_iter=1000, early_stopping=True, tol=3.0e-3, wind=6)),
            ("e", 1, dict(lbda=2.0, nb_iter=300, early_stopping=True, tol=5.0e-3, wind=4))]
    for tag, v, kw in runs:
        np.random.seed(200 + v)
        x0 = np.random.randn(T)
        np.random.seed(200 + v)
        x, z, dz, J, _, _ = quiet(ref_bs.deconv, y[v].copy(), t_r, h, **kw)
        out["x0_" + tag] = x0
        out["voxel_" + tag] = np.array(v)
        out["lbda_" + tag] = np.array(kw["lbda"])
        out["nb_iter_" + tag] = np.array(kw["nb_iter"])
        out["es_" + tag] = np.array(kw["early_stopping"])
        out["tol_" + tag] = np.array(kw.get("tol", 1.0e-6))
        out["wind_" + tag] = np.array(kw.get("wind", 6))
        out["x_" + tag], out["z_" + tag], out["dz_" + tag], out["J_" + tag] = x, z, dz, J
    save("deconv_fixed", tags=np.array([r[0] for r in runs]), **out)


def gen_loops():
    out = {}
    T, t_r, dur = 300, 1.0, 20.0
    y = gen_voxels(2, T, t_r, dur, seed0=11)
    rng = np.random.RandomState(3)
    runs = [("a", 0, 2.0, 1.7, 100, False, 1e-12), ("b", 1, 1.1, 0.5, 60, False, 1e-12),
            ("c", 0, 0.8, 1.7, 400, True, 1.0e-3), ("d", 1, 1.9, 1.0, 100, True, 1.0e-2)]
    out["y"] = y
    for tag, v, theta, lbda, n, es, tol in runs:
        h, _ = ref_hm.spm_hrf(theta, t_r, dur, False)
        Hm = ref_cv.toeplitz_from_kernel(h, T, T)
        w0 = 0.01 * rng.randn(T) if tag in ("b", "d") else np.zeros(T)
        w = ref_bs._loops_deconv(y[v].astype(np.float64), w0.copy(), Hm, float(lbda), int(n),
                                 bool(es), 4, float(tol))
        out["h_" + tag], out["w0_" + tag], out["w_" + tag] = h, w0, w
        out["par_" + tag] = np.array([v, theta, lbda, n, float(es), tol])
    save("loops_deconv", tags=np.array([r[0] for r in runs]), **out)


class _ThetaRecorder:
    """Wraps the SciPy routine the reference calls so the theta trajectory is observable."""

    def __init__(self):
        self.orig = ref_bs.fmin_l_bfgs_b
        self.thetas, self.z, self.nfev = [], [], []

    def __call__(self, *a, **k):
        res = self.orig(*a, **k)
        self.thetas.append(float(np.asarray(res[0]).reshape(-1)[0]))
        self.z.append(np.array(k["args"][0]))
        self.nfev.append(res[2]["funcalls"])
        return res


def run_bd(y, t_r, **kw):
    rec = _ThetaRecorder()
    ref_bs.fmin_l_bfgs_b = rec
    try:
        x, z, dz, h, d = ref_bs.bd(y.copy(), t_r, **kw)
    finally:
        ref_bs.fmin_l_bfgs_b = rec.orig
    return x, z, dz, h, d, rec


def gen_bd():
    out = {}
    tags = []

    def add(tag, y, t_r, **kw):
        x, z, dz, h, d, rec = run_bd(y, t_r, **kw)
        tags.append(tag)
        out["y_" + tag] = y
        out["t_r_" + tag] = np.array(t_r)
        for key in ("lbda", "hrf_dur", "nb_iter", "theta_0", "early_stopping", "wind", "tol"):
            if key in kw and kw[key] is not None:
                out[key + "_" + tag] = np.array(kw[key])
        if kw.get("z_0") is not None:
            out["z_0_" + tag] = kw["z_0"]
        out["x_" + tag], out["z_" + tag], out["dz_" + tag], out["h_" + tag] = x, z, dz, h
        out["J_" + tag], out["r_" + tag], out["g_" + tag] = d["J"], d["r"], d["g"]
        out["thetas_" + tag] = np.array(rec.thetas)
        out["nfev_" + tag] = np.array(rec.nfev)
        out["zs_" + tag] = np.array(rec.z)   # z handed to every theta step
        print(tag, "theta_end", rec.thetas[-1], "J_end", d["J"][-1], "nfev mean", np.mean(rec.nfev))

    y300 = gen_voxels(3, 300, 1.0, 20.0, seed0=21)
    add("t300_v0", y300[0], 1.0, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=100)
    add("t300_v1", y300[1], 1.0, lbda=1.0, hrf_dur=20.0, nb_iter=60)
    add("t300_es", y300[2], 1.0, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=100,
        early_stopping=True, wind=4, tol=1.0e-3)
    y240 = gen_voxels(2, 240, 0.75, 20.0, seed0=31)
    add("t240_v0", y240[0], 0.75, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=100)
    # warm start: z_0 and theta_0 (bold_signal.py:291-301)
    z0 = np.zeros(240)
    z0[40:56] = 1.0
    z0[120:136] = 1.0
    add("t240_warm", y240[1], 0.75, lbda=2.5, theta_0=1.2, z_0=z0, hrf_dur=20.0, nb_iter=40)
    y1200 = gen_voxels(1, 1200, 0.72, 20.0, seed0=41)
    add("t1200_v0", y1200[0], 0.72, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=10)
    # lambda so large that everything is thresholded: z == 0, flat theta cost
    add("t300_flat", y300[0], 1.0, lbda=1.0e6, theta_0=2.0, hrf_dur=20.0, nb_iter=5)
    save("bd", tags=np.array(tags), **out)


def gen_fit_err():
    T, t_r, dur = 300, 1.0, 20.0
    y, z_true, _ = gen_voxels(1, T, t_r, dur, seed0=51, return_truth=True)
    thetas = np.linspace(0.6, 1.9, 14)
    vals = np.array([ref_bs.hrf_fit_err(th, z_true[0], y[0], t_r, dur) for th in thetas])
    save("hrf_fit_err", y=y[0], z=z_true[0], thetas=thetas, vals=vals,
         t_r=np.array(t_r), dur=np.array(dur))


def gen_bd_t1200():
    """cfg4 shape (T = 1200, TR = 0.72, K = 28) with the reference's own nb_iter = 100 (about 25 s)."""
    out, tags = {}, ["t1200_n100"]
    y1200 = gen_voxels(2, 1200, 0.72, 20.0, seed0=43)
    x, z, dz, h, d, rec = run_bd(y1200[1], 0.72, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=100)
    tag = tags[0]
    out["y_" + tag], out["t_r_" + tag] = y1200[1], np.array(0.72)
    out["lbda_" + tag], out["theta_0_" + tag] = np.array(1.7), np.array(2.0)
    out["hrf_dur_" + tag], out["nb_iter_" + tag] = np.array(20.0), np.array(100)
    out["x_" + tag], out["z_" + tag], out["dz_" + tag], out["h_" + tag] = x, z, dz, h
    out["J_" + tag], out["r_" + tag], out["g_" + tag] = d["J"], d["r"], d["g"]
    out["thetas_" + tag] = np.array(rec.thetas)
    out["nfev_" + tag] = np.array(rec.nfev)
    print(tag, "theta_end", rec.thetas[-1], "J_end", d["J"][-1])
    save("bd_t1200", tags=np.array(tags), **out)


def gen_deconv_auto():
    """``deconv(lbda=None)`` (bold_signal.py:99-214) of the live reference.  PyWavelets is absent, so
    ``mad_daub_noise_est`` -- the name ``bold_signal`` imported -- is replaced by a function that
    returns the STORED sigma; everything else is the reference's own code."""
    out, tags = {}, []
    orig = ref_bs.mad_daub_noise_est

    def add(tag, y, t_r, h, sigma, seed, **kw):
        np.random.seed(seed)
        x0 = np.random.randn(len(y))
        np.random.seed(seed)
        ref_bs.mad_daub_noise_est = lambda x, c=0.6744: sigma
        try:
            x, z, dz, J, R, G = quiet(ref_bs.deconv, y.copy(), t_r, h, lbda=None, **kw)
        finally:
            ref_bs.mad_daub_noise_est = orig
        tags.append(tag)
        out["y_" + tag], out["h_" + tag], out["x0_" + tag] = y, h, x0
        out["t_r_" + tag], out["sigma_" + tag] = np.array(t_r), np.array(sigma)
        for key, default in (("nb_iter", 1000), ("nb_sub_iter", 1000), ("early_stopping", True),
                             ("tol", 1.0e-6), ("wind", 6)):
            out[key + "_" + tag] = np.array(kw.get(key, default))
        out["x_" + tag], out["z_" + tag], out["dz_" + tag] = x, z, dz
        out["J_" + tag], out["R_" + tag], out["G_" + tag] = np.array(J), np.array(R), np.array(G)
        print(tag, "outer iterations", len(J), "J_end", J[-1])

    y300 = gen_voxels(3, 300, 1.0, 20.0, seed0=61)
    h20, _ = ref_hm.spm_hrf(1.0, 1.0, 20.0, True)
    # a: no early stopping at all, every loop runs to its count
    add("a", y300[0], 1.0, h20, 0.35, 300, nb_iter=12, nb_sub_iter=60, early_stopping=False)
    # b: inner early stops (Q5 window) and the alpha window stop both fire
    add("b", y300[1], 1.0, h20, 0.5, 301, nb_iter=60, nb_sub_iter=300, early_stopping=True,
        tol=3.0e-2, wind=6)
    # c: another window length, tighter tolerance: inner stops, outer runs out
    add("c", y300[2], 1.0, h20, 0.2, 302, nb_iter=15, nb_sub_iter=200, early_stopping=True,
        tol=2.0e-3, wind=4)
    # d: the shape of examples/synth_data/deconv.py (cfg1: T = 600, TR = 1, HRF of 30 s at delta 1.5)
    y600 = gen_voxels(1, 600, 1.0, 30.0, seed0=71)
    h30, _ = ref_hm.spm_hrf(1.5, 1.0, 30.0, True)
    add("d", y600[0], 1.0, h30, 0.4, 303, nb_iter=40, nb_sub_iter=150, early_stopping=True,
        tol=5.0e-3, wind=6)
    # e: alpha window stop with wind = 4
    add("e", y300[1], 1.0, h20, 1.0, 304, nb_iter=60, nb_sub_iter=300, early_stopping=True,
        tol=2.0e-2, wind=4)
    save("deconv_auto", tags=np.array(tags), **out)


def gen_hrf_estim():
    """``hrf_estim`` (bold_signal.py:225-239): h and the Tracker's cost per L-BFGS-B iterate."""
    out, tags = {}, []
    for tag, T, t_r, dur, seed in (("a", 300, 1.0, 20.0, 81), ("b", 240, 0.75, 20.0, 82),
                                   ("c", 600, 1.0, 30.0, 83)):
        y, z_true, _ = gen_voxels(1, T, t_r, dur, seed0=seed, return_truth=True)
        h, J = ref_bs.hrf_estim(z_true[0], y[0], t_r, dur)
        tags.append(tag)
        out["y_" + tag], out["z_" + tag] = y[0], z_true[0]
        out["t_r_" + tag], out["dur_" + tag] = np.array(t_r), np.array(dur)
        out["h_" + tag], out["J_" + tag] = h, np.array(J, dtype=np.float64).reshape(-1)
        print(tag, "iterates", len(J), "J_end", float(np.asarray(J[-1]).reshape(-1)[0]))
    save("hrf_estim", tags=np.array(tags), **out)


HRF_PARAM_GRID = [
    dict(delta=1.0, t_r=1.0, dur=20.0, p_delay=5, undershoot=15.0),
    dict(delta=0.8, t_r=0.75, dur=25.0, p_disp=0.9, u_disp=1.2, p_u_ratio=0.3),
    dict(delta=1.4, t_r=2.0, dur=32.0, onset=0.5),        # onset / dt = 500 s: identically zero taps
    dict(delta=1.4, t_r=2.0, dur=32.0, onset=0.004),      # shifted by 4 s
    dict(delta=0.7, t_r=1.0, dur=20.0, onset=-0.002, p_delay=5.5),
    dict(delta=1.0, t_r=0.5, dur=20.0, dt=0.002),
    dict(delta=1.9, t_r=0.3, dur=18.0, p_delay=7, undershoot=12.0, p_disp=1.1, u_disp=0.8,
         p_u_ratio=0.1, onset=0.001),
]


def gen_hrf_params():
    """``spm_hrf`` with non-default shape parameters (hrf_model.py:12-39)."""
    out = {"n": np.array(len(HRF_PARAM_GRID))}
    for idx, kw in enumerate(HRF_PARAM_GRID):
        for norm, key in ((False, "h"), (True, "hn")):
            h, t = ref_hm.spm_hrf(normalized_hrf=norm, **kw)
            out["%s%d" % (key, idx)] = h
        out["t%d" % idx] = t
        out["kw%d" % idx] = np.array(sorted(kw.items()), dtype=object).astype(str)
    save("spm_hrf_params", **out)


GENERATORS = {"ops": gen_ops, "spm_hrf": gen_hrf, "lipschitz": gen_lipschitz, "deconv_fixed": gen_deconv,
              "loops_deconv": gen_loops, "hrf_fit_err": gen_fit_err, "bd": gen_bd, "bd_t1200": gen_bd_t1200,
              "deconv_auto": gen_deconv_auto, "hrf_estim": gen_hrf_estim, "spm_hrf_params": gen_hrf_params}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(GENERATORS)):
        GENERATORS[name]()
