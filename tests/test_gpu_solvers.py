"""GPU parity of the persistent solver kernels (deconv, bd, hrf_estim) through the public
API / C ABI against (a) golden vectors of the live reference and (b) the CPU oracle on the
same seeded inputs.

Stated tolerances (DESIGN.md section "Parity"):
  deconvolution stage, FP64      1e-9  relative  (north_star)
  deconvolution stage, FP32      1e-4  relative  (north_star)
  bd end to end vs the oracle running the SAME exact theta step, FP64   1e-7
  bd end to end vs the reference (SciPy L-BFGS-B theta step), FP64      theta 1.5e-6 abs, z/x/h 8e-7, J 8e-8
                                                 (T = 1200: 1e-5, 5e-6, 4e-7); <= 10x the measured error
  bd end to end, FP32 vs the reference and vs FP64   1e-4 on z, h, J, theta (north_star; measured <= 2e-5)
  deconv(lbda=None) vs the reference with sigma injected, FP64   1e-9, same outer stop iteration
  hrf_estim vs the reference   h 2e-6 (the reference's L-BFGS-B theta is ~1e-7 off the minimiser)
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pybold_oracle as orc  # noqa: E402
from pybold_b200.synth import gen_voxels  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


def test_deconv_fixed_lambda_vs_reference_golden(golden):
    import pybold_b200 as pb
    g = golden("deconv_fixed")
    for tag in g["tags"]:
        y = g["y"][int(g["voxel_" + tag])]
        x, z, dz, J, R, G = pb.deconv(y, 1.0, g["h"], lbda=float(g["lbda_" + tag]),
                                      early_stopping=bool(g["es_" + tag]), tol=float(g["tol_" + tag]),
                                      wind=int(g["wind_" + tag]), nb_iter=int(g["nb_iter_" + tag]),
                                      x0=g["x0_" + tag])
        assert R is None and G is None
        assert len(J) == len(g["J_" + tag]), tag          # same early-stop iteration (Q5)
        assert rel(dz, g["dz_" + tag]) < 1e-9, tag
        assert rel(z, g["z_" + tag]) < 1e-9, tag
        assert rel(x, g["x_" + tag]) < 1e-9, tag
        assert rel(J, g["J_" + tag]) < 1e-9, tag


def test_deconv_uses_global_rng_like_reference(golden):
    import pybold_b200 as pb
    g = golden("deconv_fixed")
    np.random.seed(200)      # the seed make_golden.py used for voxel 0 (tag "a")
    x, z, dz, J, _, _ = pb.deconv(g["y"][0], 1.0, g["h"], lbda=1.0, early_stopping=False, nb_iter=200)
    assert rel(dz, g["dz_a"]) < 1e-9


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-9), (np.float32, 1e-4)])
def test_deconv_batch_vs_oracle(dt, tol):
    """cfg2 shape (T=300, K=20, shared HRF, 200 iterations) on 48 seeded voxels."""
    import pybold_b200 as pb
    V, T = 48, 300
    y = gen_voxels(V, T, 1.0, 20.0, seed0=1000)
    h, _ = orc.spm_hrf(1.0, 1.0, 20.0, True)
    x0 = np.random.RandomState(0).randn(T)
    x, z, dz, J, _, _ = pb.deconv(y.astype(dt), 1.0, h.astype(dt), lbda=1.0, early_stopping=False,
                                  nb_iter=200, x0=x0.astype(dt))
    assert x.shape == (V, T) and J.shape == (V, 200) and x.dtype == dt
    Lc = 0.9 * orc.spectral_radius_est(orc.HrfIntegOperator(h, T), x0)
    for v in range(0, V, 5):
        xo, zo, wo, Jo, _ = orc.deconv_fixed_lbda(y[v], h, 1.0, lipschitz=Lc, early_stopping=False,
                                                  nb_iter=200)
        assert rel(dz[v], wo) < tol and rel(z[v], zo) < tol and rel(x[v], xo) < tol
        assert rel(J[v], Jo) < tol


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_deconv_early_stopping_group_kernel_with_work_queue(dt):
    """The reference's default call (early_stopping=True, bold_signal.py:13) runs on the group layout:
    every voxel stops at its own iteration, groups pull voxels from an atomic queue.  (a) iteration
    counts and results equal the oracle's (FP64); (b) a voxel's result does not depend on where it sits
    in the batch, i.e. on which group pulled it and when: permuted batch == permuted results, bit for
    bit; (c) it equals the one-voxel call."""
    import pybold_b200 as pb
    from pybold_b200 import _lib
    V, T = 333, 300
    assert _lib.lib.pb_solver_variant(T, 20, int(dt == np.float64)) // 1000000 == 16
    y = gen_voxels(V, T, 1.0, 20.0, seed0=1500)
    y *= np.linspace(0.2, 3.0, V)[:, None]           # spread the stopping iterations
    h, _ = orc.spm_hrf(1.0, 1.0, 20.0, True)
    x0 = np.random.RandomState(0).randn(T)
    kw = dict(lbda=1.0, early_stopping=True, tol=5e-3, wind=6, nb_iter=400, x0=x0.astype(dt))
    x, z, dz, J, _, _ = pb.deconv(y.astype(dt), 1.0, h.astype(dt), **kw)
    n_it = np.sum(~np.isnan(J), axis=1)
    assert n_it.min() < n_it.max() and n_it.min() > 6 and n_it.max() <= 400
    perm = np.random.RandomState(1).permutation(V)
    xp, zp, dzp, Jp, _, _ = pb.deconv(y[perm].astype(dt), 1.0, h.astype(dt), **kw)
    assert np.array_equal(zp, z[perm]) and np.array_equal(dzp, dz[perm]) and np.array_equal(xp, x[perm])
    assert np.array_equal(np.isnan(Jp), np.isnan(J[perm]))
    assert np.array_equal(np.nan_to_num(Jp), np.nan_to_num(J[perm]))
    for v in (0, 7, V - 1):
        x1, z1, dz1, J1, _, _ = pb.deconv(y[v].astype(dt), 1.0, h.astype(dt), **kw)
        assert len(J1) == n_it[v] and np.array_equal(z1, z[v]) and np.array_equal(J1, J[v, :n_it[v]])
    if dt == np.float64:
        Lc = 0.9 * orc.spectral_radius_est(orc.HrfIntegOperator(h, T), x0)
        for v in range(0, V, 37):
            xo, zo, wo, Jo, n_o = orc.deconv_fixed_lbda(y[v], h, 1.0, lipschitz=Lc, early_stopping=True,
                                                        tol=5e-3, wind=6, nb_iter=400)
            assert n_o == n_it[v], v
            assert rel(dz[v], wo) < 1e-9 and rel(x[v], xo) < 1e-9
            assert rel(J[v, :n_o], np.asarray(Jo)) < 1e-9


@pytest.mark.parametrize("T", [3300, 4096])
def test_deconv_early_stopping_long_series_fp64(T):
    """FP64 deconv with the reference's default early stopping on series beyond ~3200 scans: the generic
    kernel's ring of past iterates no longer fits in shared memory next to its work rows, the oldest rows
    then live in the voxel's own output rows until the final store.  Iteration count and results as the
    oracle's."""
    import pybold_b200 as pb
    y = gen_voxels(2, T, 1.0, 20.0, seed0=8800 + T)
    h, _ = orc.spm_hrf(1.0, 1.0, 20.0, True)
    x0 = np.random.RandomState(3).randn(T)
    kw = dict(lbda=1.0, early_stopping=True, tol=5e-2, wind=6, nb_iter=60)
    x, z, dz, J, _, _ = pb.deconv(y, 1.0, h, x0=x0, **kw)
    n_it = np.sum(~np.isnan(J), axis=1)
    Lc = 0.9 * orc.spectral_radius_est(orc.HrfIntegOperator(h, T), x0)
    for v in range(2):
        xo, zo, wo, Jo, n_o = orc.deconv_fixed_lbda(y[v], h, 1.0, lipschitz=Lc, early_stopping=True, tol=5e-2,
                                                    wind=6, nb_iter=60)
        assert n_it[v] == n_o and 7 < n_o < 60
        assert rel(z[v], zo) < 1e-9 and rel(x[v], xo) < 1e-9 and rel(J[v, :n_o], Jo) < 1e-9
    # without early stopping the same shapes run as before
    x2, z2, dz2, J2, _, _ = pb.deconv(y, 1.0, h, x0=x0, lbda=1.0, early_stopping=False, nb_iter=12)
    assert J2.shape[1] == 12 and np.all(np.isfinite(z2))


def test_deconv_ragged_and_edge_shapes():
    import pybold_b200 as pb
    rng = np.random.RandomState(3)
    for (T, K, n) in [(1, 1, 3), (5, 9, 10), (33, 20, 25), (301, 27, 30), (1200, 28, 12)]:
        y = rng.randn(2, T)
        h = np.abs(rng.randn(K)) + 0.1
        x0 = rng.randn(T)
        x, z, dz, J, _, _ = pb.deconv(y, 1.0, h, lbda=0.05, early_stopping=False, nb_iter=n, x0=x0)
        Lc = 0.9 * orc.spectral_radius_est(orc.HrfIntegOperator(h, T), x0)
        xo, zo, wo, Jo, _ = orc.deconv_fixed_lbda(y[1], h, 0.05, lipschitz=Lc, early_stopping=False,
                                                  nb_iter=n)
        assert rel(dz[1], wo) < 1e-9 and rel(J[1], Jo) < 1e-9, (T, K)
    # empty batch
    out = pb.deconv(np.zeros((0, 16)), 1.0, np.ones(4), lbda=1.0, nb_iter=3, x0=np.ones(16))
    assert out[0].shape == (0, 16)


def _bd_kwargs(g, tag):
    kw = {}
    for key in ("lbda", "hrf_dur", "nb_iter", "theta_0", "early_stopping", "wind", "tol"):
        if key + "_" + tag in g.files:
            kw[key] = g[key + "_" + tag].item()
    if "z_0_" + tag in g.files:
        kw["z_0"] = g["z_0_" + tag]
    return kw


# Tolerances of the bd-vs-reference gates: <= 10x what the exact theta step measures against the
# reference's L-BFGS-B trajectory on the CPU (tests/test_oracle_golden.py: theta 1.4e-7, z/h 8.5e-8,
# J 7.7e-9 at T <= 300; theta 1e-6, z 3e-7, h 5.7e-7, J 3.6e-8 at T = 1200).
def _bd_gate(tag):
    if tag.startswith("t1200"):
        return dict(theta=1e-5, sig=5e-6, J=4e-7, g=3e-6)
    return dict(theta=1.5e-6, sig=8e-7, J=8e-8, g=5e-7)


def _check_bd_vs_golden(g, tag, got, gate, theta_key="thetas_"):
    x, z, dz, h, d = got
    assert len(d["J"]) == len(g["J_" + tag])
    assert abs(d["theta"] - g[theta_key + tag][-1]) < gate["theta"]
    assert rel(z, g["z_" + tag]) < gate["sig"]
    assert rel(x, g["x_" + tag]) < gate["sig"]
    assert rel(dz, g["dz_" + tag]) < gate["sig"]
    assert rel(h, g["h_" + tag]) < gate["sig"]
    assert rel(d["J"], g["J_" + tag]) < gate["J"]
    assert rel(d["r"], g["r_" + tag]) < gate["J"]
    assert rel(d["g"], g["g_" + tag]) < gate["g"]
    assert d["l_alpha"] == []


@pytest.mark.parametrize("tag", ["t300_v0", "t300_v1", "t240_v0", "t240_warm", "t1200_v0",
                                 "t300_flat", "t300_es"])
def test_bd_vs_reference_golden(golden, tag):
    import pybold_b200 as pb
    g = golden("bd")
    got = pb.bd(g["y_" + tag], float(g["t_r_" + tag]), **_bd_kwargs(g, tag))
    _check_bd_vs_golden(g, tag, got, _bd_gate(tag))


def test_bd_t1200_full_iterations_vs_reference_golden(golden):
    """cfg4 shape with the reference's own nb_iter = 100 (10 000 inner iterations per outer pass)."""
    import pybold_b200 as pb
    g = golden("bd_t1200")
    tag = "t1200_n100"
    got = pb.bd(g["y_" + tag], float(g["t_r_" + tag]), **_bd_kwargs(g, tag))
    _check_bd_vs_golden(g, tag, got, _bd_gate(tag))


@pytest.mark.parametrize("name,tag", [("bd", "t300_v0"), ("bd", "t240_v0"), ("bd", "t300_v1"),
                                      ("bd_t1200", "t1200_n100")])
def test_bd_fp32_vs_reference_golden(golden, name, tag):
    """north_star's FP32 gate (1e-4 relative on the neural signal, theta and the objective trace)
    against the REFERENCE's float64 output, not against our own FP64 build."""
    import pybold_b200 as pb
    g = golden(name)
    x, z, dz, h, d = pb.bd(g["y_" + tag].astype(np.float32), float(g["t_r_" + tag]), **_bd_kwargs(g, tag))
    assert z.dtype == np.float32
    assert abs(d["theta"] - g["thetas_" + tag][-1]) < 1e-4
    assert rel(z, g["z_" + tag]) < 1e-4 and rel(x, g["x_" + tag]) < 1e-4
    assert rel(h, g["h_" + tag]) < 1e-4
    assert rel(d["J"], g["J_" + tag]) < 1e-4 and rel(d["r"], g["r_" + tag]) < 1e-4


@pytest.mark.parametrize("T,t_r,V", [(300, 1.0, 6), (240, 0.75, 4)])
def test_bd_fp64_vs_oracle_same_theta_algorithm(T, t_r, V):
    """Device bd against the oracle running the same exact theta step: stage-exact parity."""
    import pybold_b200 as pb
    y = gen_voxels(V, T, t_r, 20.0, seed0=2000)
    x, z, dz, h, d = pb.bd(y, t_r, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=40)
    for v in range(V):
        xo, zo, wo, ho, do = orc.bd(y[v], t_r, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=40,
                                    theta_solver="exact")
        assert rel(dz[v], wo) < 1e-7 and rel(z[v], zo) < 1e-7 and rel(x[v], xo) < 1e-7
        assert rel(h[v], ho) < 1e-7
        assert rel(d["J"][v], do["J"]) < 1e-8 and rel(d["r"][v], do["r"]) < 1e-8
        assert rel(d["g"][v], do["g"]) < 1e-7


def test_bd_fp32_vs_fp64():
    import pybold_b200 as pb
    V, T = 64, 300
    y = gen_voxels(V, T, 1.0, 20.0, seed0=3000)
    ref = pb.bd(y, 1.0, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=100)
    got = pb.bd(y.astype(np.float32), 1.0, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=100)
    assert got[0].dtype == np.float32
    assert np.max(np.abs(got[4]["theta"] - ref[4]["theta"])) < 1e-4
    assert rel(got[4]["J"], ref[4]["J"]) < 1e-4
    assert rel(got[1], ref[1]) < 1e-4          # z  (measured 6e-6, profiles/r01_fp32_accuracy.txt)
    assert rel(got[3], ref[3]) < 1e-4          # h


@pytest.mark.parametrize("T,t_r", [(300, 1.0), (240, 0.75), (150, 1.0)])
def test_bd_early_stopping_group_kernel(T, t_r):
    """bd(early_stopping=True) on the group layout (round 2): the inner Q6 stop fires at a different iteration
    for every voxel of a warp and the outer Q7 stop ends the run early.  FP64 against the oracle running the
    same exact theta step; a voxel solved in a batch == the same voxel solved alone (bit exact); FP32 1e-4."""
    import pybold_b200 as pb
    from pybold_b200 import _lib
    V = 5
    y = gen_voxels(V, T, t_r, 20.0, seed0=2500 + T)
    y *= np.linspace(0.5, 2.0, V)[:, None]
    kw = dict(lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=30, early_stopping=True, wind=4, tol=2.0e-3)
    x, z, dz, h, d = pb.bd(y, t_r, **kw)
    n_tr = d["n_trace"]
    assert np.all(n_tr < 32) and np.all(n_tr >= 6)            # Q7 ended every run early
    for v in range(V):
        xo, zo, wo, ho, do = orc.bd(y[v], t_r, theta_solver="exact", **kw)
        assert len(do["J"]) == n_tr[v], v
        assert rel(z[v], zo) < 1e-7 and rel(dz[v], wo) < 1e-7 and rel(h[v], ho) < 1e-7, v
        assert rel(d["J"][v, :n_tr[v]], do["J"]) < 1e-8, v
        x1, z1, dz1, h1, d1 = pb.bd(y[v], t_r, **kw)
        assert np.array_equal(z1, z[v]) and np.array_equal(d1["J"], d["J"][v, :n_tr[v]]), v
    K = h.shape[1]
    if _lib.lib.pb_solver_variant(T, K, 0) // 1000000 <= 16:
        x32, z32, dz32, h32, d32 = pb.bd(y.astype(np.float32), t_r, **kw)
        assert np.array_equal(d32["n_trace"], n_tr)
        assert rel(z32, z) < 1e-4 and rel(h32, h) < 1e-4
        for v in range(V):
            assert rel(d32["J"][v, :n_tr[v]], d["J"][v, :n_tr[v]]) < 1e-4


def test_bd_per_voxel_parameters_and_tensor_io():
    import pybold_b200 as pb
    V, T = 5, 300
    y = gen_voxels(V, T, 1.0, 20.0, seed0=4000)
    lb = np.array([0.5, 1.0, 1.7, 2.5, 4.0])
    th0 = np.array([2.0, 1.5, 1.0, 0.8, 2.0])
    yt = torch.as_tensor(y, device="cuda")
    x, z, dz, h, d = pb.bd(yt, 1.0, lbda=lb, theta_0=th0, nb_iter=20)
    assert isinstance(x, torch.Tensor) and x.is_cuda and d["J"].shape == (V, 22)
    for v in (0, 3):
        x1, z1, dz1, h1, d1 = pb.bd(y[v], 1.0, lbda=lb[v], theta_0=th0[v], nb_iter=20)
        assert rel(z[v].cpu().numpy(), z1) < 1e-12
        assert rel(d["J"][v].cpu().numpy(), d1["J"]) < 1e-12
    with pytest.raises(ValueError):
        pb.bd(y, 1.0, theta_0=2.5)


def test_hrf_estim_vs_exact_minimiser(golden):
    import pybold_b200 as pb
    g = golden("hrf_fit_err")
    y, z = g["y"], g["z"]
    h, J = pb.hrf_estim(z, y, 1.0, 20.0)
    th = orc.theta_step_exact(2.0, z, y, 1.0, 20.0, [(0.6, 1.9)])
    assert rel(h, orc.spm_hrf_closed_form(th, 1.0, 20.0)) < 1e-9
    assert abs(J[-1] / orc.hrf_fit_err_fast(th, z, y, 1.0, 20.0) - 1) < 1e-9
    for t, val in zip(g["thetas"][::4], g["vals"][::4]):
        assert abs(pb.hrf_fit_err(t, z, y, 1.0, 20.0) / val - 1) < 1e-12


@pytest.mark.parametrize("T,t_r", [(296, 1.0), (299, 1.0), (304, 1.0), (233, 0.75), (305, 1.0), (64, 1.0),
                                   (600, 1.0), (585, 1.0), (640, 0.75), (1195, 0.72), (1280, 0.72),
                                   (284, 0.7535), (200, 1.0), (209, 0.72), (320, 0.72), (400, 1.0),
                                   (500, 0.72), (700, 0.72), (900, 0.72), (350, 1.0), (330, 0.72), (340, 1.0), (360, 1.0), (650, 1.0), (800, 0.72), (150, 1.0), (100, 0.72), (190, 1.0), (128, 2.0),
                                   (1050, 0.72)])
def test_bd_kernel_variants_tail_shapes(T, t_r):
    """Shapes at the edges of the register-tiled variants (group kernel tail, CTA kernel with one
    and two warps per voxel, linear kernel, short series), odd batch so that one group of the last
    warp idles."""
    import pybold_b200 as pb
    from pybold_b200 import _lib
    V = 3
    n_it = 15 if T < 650 else 6
    y = gen_voxels(V, T, t_r, 20.0, seed0=5000 + T)
    x, z, dz, h, d = pb.bd(y, t_r, lbda=1.2, theta_0=2.0, hrf_dur=20.0, nb_iter=n_it)
    assert _lib.lib.pb_solver_variant(T, h.shape[1], 0) != 0      # FP32 has a register-tiled variant
    y32 = y.astype(np.float32)
    x32, z32, dz32, h32, d32 = pb.bd(y32, t_r, lbda=1.2, theta_0=2.0, hrf_dur=20.0, nb_iter=n_it)
    assert rel(z32, z) < 1e-4 and rel(h32, h) < 1e-4 and rel(d32["J"], d["J"]) < 1e-4
    for v in range(V):
        if T > 650 and v > 0:
            continue            # the dense T^3 oracle is slow at this size
        xo, zo, wo, ho, do = orc.bd(y[v], t_r, lbda=1.2, theta_0=2.0, hrf_dur=20.0, nb_iter=n_it,
                                    theta_solver="exact")
        assert rel(z[v], zo) < 1e-8 and rel(x[v], xo) < 1e-8 and rel(dz[v], wo) < 1e-8, (T, v)
        assert rel(h[v], ho) < 1e-8 and rel(d["J"][v], do["J"]) < 1e-9, (T, v)


def test_bd_generic_kernel_matches_fast_kernel():
    """K = 40 taps (t_r = 0.5 s) has no register-tiled instantiation: generic kernel."""
    import pybold_b200 as pb
    from pybold_b200 import _lib
    T, t_r = 200, 0.5
    y = gen_voxels(2, T, t_r, 20.0, seed0=6000)
    x, z, dz, h, d = pb.bd(y, t_r, lbda=1.0, nb_iter=12)
    assert h.shape[1] == 40 and _lib.lib.pb_solver_variant(T, 40, 1) == 0
    xo, zo, wo, ho, do = orc.bd(y[1], t_r, lbda=1.0, nb_iter=12, theta_solver="exact")
    assert rel(z[1], zo) < 1e-8 and rel(h[1], ho) < 1e-8 and rel(d["J"][1], do["J"]) < 1e-9


def test_bd_long_series_four_warp_variant_vs_generic_fp64():
    """1280 < T <= 2560 runs on four warps per voxel in FP32; the FP64 build has no register-tiled
    variant there (generic shared-memory kernel): two independent implementations, 1e-4."""
    import pybold_b200 as pb
    from pybold_b200 import _lib
    for T, t_r in ((2000, 0.72), (1300, 1.0), (2560, 1.0)):
        K = pb.hrf_model.hrf_len(t_r, 20.0)
        assert _lib.lib.pb_solver_variant(T, K, 0) // 1000000 in (96, 128) and _lib.lib.pb_solver_variant(T, K, 1) == 0
        y = gen_voxels(3, T, t_r, 20.0, seed0=7000 + T)
        x, z, dz, h, d = pb.bd(y, t_r, lbda=1.4, theta_0=2.0, hrf_dur=20.0, nb_iter=8)
        x32, z32, dz32, h32, d32 = pb.bd(y.astype(np.float32), t_r, lbda=1.4, theta_0=2.0, hrf_dur=20.0,
                                         nb_iter=8)
        assert rel(z32, z) < 1e-4 and rel(x32, x) < 1e-4 and rel(h32, h) < 1e-4
        assert rel(d32["J"], d["J"]) < 1e-4 and np.max(np.abs(d32["theta"] - d["theta"])) < 1e-4


@pytest.mark.parametrize("T,t_r,n_it", [(200, 0.5, 12), (600, 0.5, 8), (90, 0.5, 10), (40, 0.5, 10), (1000, 0.5, 6),
                                        (2000, 0.5, 4), (3000, 0.5, 3), (300, 0.32, 10), (1200, 0.32, 5),
                                        (4096, 0.32, 3), (3000, 1.0, 3), (4096, 0.72, 3), (2600, 1.0, 4),
                                        (3500, 1.0, 3), (3600, 0.5, 3), (3900, 0.72, 3),
                                        # gap-filling variants of the 40- and 64-tap families
                                        (450, 0.5, 8), (700, 0.5, 8), (900, 0.5, 6), (1500, 0.5, 5), (1800, 0.5, 4),
                                        (150, 0.32, 10), (200, 0.32, 10), (450, 0.32, 8), (700, 0.32, 8),
                                        (900, 0.32, 6), (1500, 0.32, 5)])
def test_bd_short_tr_and_long_series_variants(T, t_r, n_it):
    """Round 2 variants: 28 < K <= 64 taps (TR 0.5 s -> K = 40, TR 0.32 s -> K = 63) and 2560 < T <= 4096
    (six or eight warps per voxel) run register-tiled in FP32; the FP64 build of these shapes is the generic
    shared-memory kernel: two independent implementations, north_star's 1e-4."""
    import pybold_b200 as pb
    from pybold_b200 import _lib
    K = pb.hrf_model.hrf_len(t_r, 20.0)
    assert _lib.lib.pb_solver_variant(T, K, 0) != 0
    y = gen_voxels(3, T, t_r, 20.0, seed0=7500 + T)
    x, z, dz, h, d = pb.bd(y, t_r, lbda=1.4, theta_0=2.0, hrf_dur=20.0, nb_iter=n_it)
    x32, z32, dz32, h32, d32 = pb.bd(y.astype(np.float32), t_r, lbda=1.4, theta_0=2.0, hrf_dur=20.0,
                                     nb_iter=n_it)
    assert h32.shape == (3, K)
    # north_star's 1e-4 is gated at the BASELINE shapes (cfg3, cfg4: golden tests above) and holds up to
    # T ~ 1300 (measured here: 2e-5 at T = 1000, 6e-5 at T = 1200 with 63 taps).  Beyond, FP32 and FP64
    # drift apart with the series length: measured z 1.2e-4, theta 1.4e-4 at T = 2000; z 3.6e-4 at 3000;
    # z 6e-4, h 1.2e-3, theta 7.5e-4 at 4096.  It is the theta step's conditioning, not the arithmetic of
    # the inner loops: with theta held fixed the same runs agree to 4e-6 (T = 2000) and 7e-6 (T = 4096), and
    # accumulating the theta moments in double changes nothing (tools/debug_case.py).  Those shapes lie
    # outside BASELINE.json; they are gated at the stated looser bounds and documented.
    tol = 1e-4 if T <= 1300 else (4e-4 if T <= 2600 else 2.5e-3)
    assert rel(z32, z) < tol and rel(x32, x) < tol and rel(h32, h) < tol
    assert rel(d32["J"], d["J"]) < 1e-4 and np.max(np.abs(d32["theta"] - d["theta"])) < tol
    if T <= 300:
        xo, zo, wo, ho, do = orc.bd(y[1], t_r, lbda=1.4, theta_0=2.0, hrf_dur=20.0, nb_iter=n_it,
                                    theta_solver="exact")
        assert rel(z32[1], zo) < 1e-4 and rel(h32[1], ho) < 1e-4 and rel(d32["J"][1], do["J"]) < 1e-4


def test_bd_streamed_host_batch_equals_single_launch():
    """Host batches above the streaming threshold are solved chunk by chunk with overlapped copies;
    the result must be bit-identical to the single-launch path, including per-voxel parameters
    and a warm start."""
    import pybold_b200 as pb
    from pybold_b200 import bold_signal as bs
    V, T = 9001, 48
    rng = np.random.RandomState(1)
    y = gen_voxels(64, T, 1.0, 20.0, seed0=7000)
    y = np.tile(y, (V // 64 + 1, 1))[:V] * rng.uniform(0.5, 1.5, (V, 1))
    lb = rng.uniform(0.5, 2.0, V)
    z0 = np.zeros((V, T))
    z0[:, 10:20] = 1.0
    old = bs._STREAM_TARGET_CHUNK
    bs._STREAM_TARGET_CHUNK = 4096      # force 3 chunks
    try:
        a = pb.bd(y.astype(np.float32), 1.0, lbda=lb, z_0=z0, nb_iter=6)
    finally:
        bs._STREAM_TARGET_CHUNK = old
    yt = torch.as_tensor(y.astype(np.float32), device="cuda")
    b = pb.bd(yt, 1.0, lbda=lb, z_0=z0, nb_iter=6)
    assert isinstance(a[0], np.ndarray) and a[0].shape == (V, T)
    for u, v in zip(a[:4], b[:4]):
        assert np.array_equal(u, v.cpu().numpy())
    for k in ("J", "r", "g", "theta", "n_trace"):
        assert np.array_equal(a[4][k], b[4][k].cpu().numpy()), k
    # results too large for pinned host tensors go through the two pinned staging slots into pageable
    # memory: same bits (forced here by a zero limit); a pinned input tensor is uploaded in place
    old_lim = bs._PINNED_RESULT_LIMIT
    bs._STREAM_TARGET_CHUNK, bs._PINNED_RESULT_LIMIT = 4096, 0
    try:
        c = pb.bd(torch.from_numpy(y.astype(np.float32)).pin_memory(), 1.0, lbda=lb, z_0=z0, nb_iter=6)
    finally:
        bs._STREAM_TARGET_CHUNK, bs._PINNED_RESULT_LIMIT = old, old_lim
    assert isinstance(c[0], torch.Tensor) and not c[0].is_cuda and not c[0].is_pinned()
    for u, v in zip(a[:4], c[:4]):
        assert np.array_equal(u, v.numpy())
    for k in ("J", "r", "g", "theta", "n_trace"):
        assert np.array_equal(a[4][k], c[4][k].numpy()), k


def test_deconv_auto_lambda_vs_oracle():
    """deconv(lbda=None) (row N1): noise-constrained lambda loop with sigma supplied, against the
    oracle restatement of pybold/bold_signal.py:99-214 (survey: matches the reference at 1e-16)."""
    import pybold_b200 as pb
    T = 200
    y = gen_voxels(2, T, 1.0, 20.0, seed0=8000)
    h, _ = orc.spm_hrf(1.2, 1.0, 20.0, True)
    x0 = np.random.RandomState(1).randn(T)
    sigma = 0.3
    for es in (False, True):
        x, z, dz, J, R, G = pb.deconv(y[0], 1.0, h, lbda=None, nb_iter=8, nb_sub_iter=40,
                                      early_stopping=es, tol=1e-3, x0=x0, sigma=sigma)
        xo, zo, wo, Jo, Ro, Go, _ = orc.deconv_auto_lbda(y[0], h, sigma, x0_power=x0, early_stopping=es,
                                                         tol=1e-3, nb_iter=8, nb_sub_iter=40)
        assert len(J) == len(Jo) and isinstance(J, list)
        assert rel(dz, wo) < 1e-9 and rel(z, zo) < 1e-9 and rel(x, xo) < 1e-9
        assert rel(J, Jo) < 1e-9 and rel(R, Ro) < 1e-9 and rel(G, Go) < 1e-9
    # batched: both voxels at once equal the one-by-one runs
    xb, zb, dzb, Jb, Rb, Gb = pb.deconv(y, 1.0, h, lbda=None, nb_iter=5, nb_sub_iter=30,
                                        early_stopping=False, x0=x0, sigma=sigma)
    x1, z1, dz1, J1, R1, G1 = pb.deconv(y[1], 1.0, h, lbda=None, nb_iter=5, nb_sub_iter=30,
                                        early_stopping=False, x0=x0, sigma=sigma)
    assert rel(zb[1], z1) < 1e-12 and rel(Jb[1], J1) < 1e-12


def test_noise_estimate_kernel():
    """db3 MAD sigma on the device (`pb_mad_daub_noise_est_*`) == the oracle's statement of
    pybold/utils.py:10-25 (exact order statistics: only the six-tap sums may differ by rounding),
    for the reference's own call patterns: 1-D NumPy -> float, [V, T] -> one value per voxel,
    odd / even / very short series, torch in -> torch out."""
    from pybold_b200.utils import mad, mad_daub_noise_est
    for T in (301, 300, 11, 10, 9, 5, 1200, 4096):
        y = gen_voxels(5, max(T, 32), 1.0, 20.0, seed0=8100 + T)[:, :T].copy()
        want = np.array([orc.mad_daub_noise_est(v) for v in y])
        got = mad_daub_noise_est(y)
        assert isinstance(got, np.ndarray) and got.shape == (5,)
        assert np.max(np.abs(got / want - 1)) < 1e-12, T
        s1 = mad_daub_noise_est(y[2])                       # what bold_signal.py:103 does
        assert isinstance(s1, float) and abs(s1 / want[2] - 1) < 1e-12
        gt = mad_daub_noise_est(torch.as_tensor(y, device="cuda"))
        assert isinstance(gt, torch.Tensor) and gt.is_cuda and np.array_equal(gt.cpu().numpy(), got)
        g32 = mad_daub_noise_est(y.astype(np.float32))
        assert g32.dtype == np.float32 and np.max(np.abs(g32 / want - 1)) < 1e-5
    rng = np.random.RandomState(5)
    for n in (1, 2, 7, 8, 301):
        x = rng.randn(3, n)
        x[1, : n // 2] = x[1, 0]                            # ties
        want = np.array([orc.mad(v) for v in x])
        assert np.array_equal(mad(x), want), n              # exact selection: bit equal
        assert mad(x[0]) == want[0] and isinstance(mad(x[0]), float)
        assert mad(x[2], c=1.0) == orc.mad(x[2], c=1.0)


def test_deconv_auto_lambda_vs_reference_golden(golden):
    """``deconv(lbda=None)`` against the LIVE REFERENCE (sigma injected, tests/golden/make_golden.py):
    same number of outer iterations (alpha-window stop), 1e-9 on everything."""
    import pybold_b200 as pb
    g = golden("deconv_auto")
    for tag in g["tags"]:
        x, z, dz, J, R, G = pb.deconv(
            g["y_" + tag], float(g["t_r_" + tag]), g["h_" + tag], lbda=None,
            early_stopping=bool(g["early_stopping_" + tag]), tol=float(g["tol_" + tag]),
            wind=int(g["wind_" + tag]), nb_iter=int(g["nb_iter_" + tag]),
            nb_sub_iter=int(g["nb_sub_iter_" + tag]), x0=g["x0_" + tag], sigma=float(g["sigma_" + tag]))
        assert isinstance(J, list) and len(J) == len(g["J_" + tag]), tag
        assert rel(dz, g["dz_" + tag]) < 1e-9 and rel(z, g["z_" + tag]) < 1e-9, tag
        assert rel(x, g["x_" + tag]) < 1e-9, tag
        assert rel(J, g["J_" + tag]) < 1e-9 and rel(R, g["R_" + tag]) < 1e-9, tag
        assert rel(G, g["G_" + tag]) < 1e-9, tag
    # batched: every voxel gets its own outer stop iteration (masked inner solves once a voxel has stopped)
    # and the same numbers as when it is solved alone
    tag = "b"
    kw = dict(lbda=None, early_stopping=True, tol=float(g["tol_b"]), wind=int(g["wind_b"]),
              nb_iter=int(g["nb_iter_b"]), nb_sub_iter=int(g["nb_sub_iter_b"]), x0=g["x0_b"])
    sig = np.array([0.5, 0.25, 1.0, 0.5])
    yb = np.stack([g["y_b"], g["y_b"], g["y_b"], g["y_a"]])
    xb, zb, dzb, Jb, Rb, Gb = pb.deconv(yb, 1.0, g["h_b"], sigma=sig, **kw)
    n_stop = []
    for v in range(4):
        x1, z1, dz1, J1, R1, G1 = pb.deconv(yb[v], 1.0, g["h_b"], sigma=float(sig[v]), **kw)
        n_stop.append(len(J1))
        assert np.array_equal(zb[v], z1) and np.array_equal(dzb[v], dz1) and np.array_equal(xb[v], x1), v
        assert np.array_equal(Jb[v, :len(J1)], np.array(J1)) and np.all(np.isnan(Jb[v, len(J1):])), v
        assert np.array_equal(Rb[v, :len(J1)], np.array(R1)) and np.array_equal(Gb[v, :len(J1)], np.array(G1)), v
    assert n_stop[0] == len(g["J_b"]) and len(set(n_stop)) > 1 and Jb.shape[1] == max(n_stop)


def test_hrf_estim_vs_reference_golden(golden):
    """``hrf_estim`` (bold_signal.py:225-239) against the live reference: the reference's L-BFGS-B answer
    is within ~1e-7 of the exact minimiser the device finds (measured on the CPU: theta 8e-8, h 2.2e-7)."""
    import pybold_b200 as pb
    g = golden("hrf_estim")
    for tag in g["tags"]:
        h, J = pb.hrf_estim(g["z_" + tag], g["y_" + tag], float(g["t_r_" + tag]), float(g["dur_" + tag]))
        assert rel(h, g["h_" + tag]) < 2e-6, tag
        assert abs(J[-1] / g["J_" + tag][-1] - 1) < 1e-11, tag      # the cost is flat at its minimum
        assert J[0] >= J[-1]
        he, Je, the = orc.hrf_estim_exact(g["z_" + tag], g["y_" + tag], float(g["t_r_" + tag]),
                                          float(g["dur_" + tag]))
        assert rel(h, he) < 1e-9 and abs(J[0] / Je[0] - 1) < 1e-11, tag


def test_inner_loop_stage_vs_reference_numba_golden(golden):
    """Stage-wise gate of the 1e-9 FP64 target: the reference's own Numba `_loops_deconv`
    (pybold/bold_signal.py:242-278) on given (h, warm start) against the device recursion fed with the
    device Frobenius Lipschitz constant -- no theta step involved."""
    from pybold_b200 import _lib
    from pybold_b200.bold_signal import deconv_batch
    g = golden("loops_deconv")
    for tag in ("a", "b"):                       # the runs without early stopping
        v, _, lbda, n, es, _ = g["par_" + tag]
        assert not es
        y = torch.as_tensor(g["y"][int(v)], device="cuda").reshape(1, -1)
        h = torch.as_tensor(g["h_" + tag], device="cuda")
        w0 = torch.as_tensor(g["w0_" + tag], device="cuda").reshape(1, -1).contiguous()
        L = torch.empty(1, dtype=torch.float64, device="cuda")
        assert _lib.lib.pb_lipschitz_frob_f64(h.data_ptr(), 0, L.data_ptr(), 1, y.shape[1], h.numel(), 0) == 0
        x, z, dz, J, n_it = deconv_batch(y, h, float(lbda), L, w0, False, 1e-6, 6, int(n))
        assert rel(dz[0].cpu().numpy(), g["w_" + tag]) < 1e-9, tag


def test_deconv_lbda_path_matches_single_lambda_calls():
    """cfg5-style regularisation path: batched over (lambda, voxel) == one deconv per lambda."""
    import pybold_b200 as pb
    from pybold_b200.bold_signal import deconv_lbda_path
    V, T = 7, 600
    y = gen_voxels(V, T, 1.0, 20.0, seed0=9000)
    h, _ = orc.spm_hrf(1.0, 1.0, 20.0, True)
    x0 = np.random.RandomState(2).randn(T)
    lbdas = np.geomspace(0.05, 20, 5)
    x, z, dz, J = deconv_lbda_path(y, 1.0, h, lbdas, nb_iter=40, x0=x0, max_problems=16)
    assert z.shape == (5, V, T) and J.shape == (5, V, 40)
    for i in (0, 4):
        x1, z1, dz1, J1, _, _ = pb.deconv(y, 1.0, h, lbda=float(lbdas[i]), early_stopping=False,
                                         nb_iter=40, x0=x0)
        assert np.array_equal(z[i], z1) and np.array_equal(J[i], J1)
    Lc = 0.9 * orc.spectral_radius_est(orc.HrfIntegOperator(h, T), x0)
    xo, zo, wo, Jo, _ = orc.deconv_fixed_lbda(y[3], h, lbdas[2], lipschitz=Lc, early_stopping=False, nb_iter=40)
    assert rel(dz[2, 3], wo) < 1e-9 and rel(J[2, 3], Jo) < 1e-9


@pytest.mark.parametrize("T,t_r", [(300, 1.0), (1200, 0.72), (150, 1.0), (333, 1.0)])
def test_bd_degenerate_inputs_do_not_hang_or_leak(T, t_r):
    """All-zero, constant, huge, tiny and NaN voxels in one batch: the call returns, finite voxels give
    finite results identical to a batch without the bad voxels, the NaN stays in its own voxel."""
    import pybold_b200 as pb
    rng = np.random.RandomState(3)
    good = gen_voxels(6, T, t_r, 20.0, seed0=9000 + T).astype(np.float32)
    y = np.concatenate([good, np.zeros((1, T), np.float32), np.full((1, T), 3.5, np.float32),
                        (1e18 * good[:1]), (1e-18 * good[1:2]), good[2:3].copy()], axis=0)
    y[-1, T // 2] = np.nan
    lb = np.array([1.7] * 6 + [1.7, 1.7, 1.7, 1.7, 1.7], dtype=np.float32)
    x, z, dz, h, d = pb.bd(y, t_r, lbda=lb, theta_0=2.0, hrf_dur=20.0, nb_iter=12)
    xg, zg, dzg, hg, dg = pb.bd(good, t_r, lbda=1.7, theta_0=2.0, hrf_dur=20.0, nb_iter=12)
    assert np.array_equal(z[:6], zg) and np.array_equal(h[:6], hg) and np.array_equal(d["J"][:6], dg["J"])
    assert np.all(z[6] == 0) and np.all(np.isfinite(h[6])) and 0.6 <= d["theta"][6] <= 1.9   # flat cost: theta stays
    for v in (7, 9):
        assert np.all(np.isfinite(z[v])) and np.all(np.isfinite(h[v])) and np.all(np.isfinite(d["J"][v]))
    assert np.all(np.isfinite(h[8])) and 0.6 <= d["theta"][8] <= 1.9
    assert np.all(np.isfinite(h[10]))                      # taps stay finite, the estimate itself is NaN
    assert np.isnan(z[10]).any() and not np.isnan(z[:10]).any()
    # lbda = 0 and a huge lbda
    x0, z0, _, _, d0 = pb.bd(good[:2], t_r, lbda=0.0, nb_iter=8)
    xh, zh, _, _, dh = pb.bd(good[:2], t_r, lbda=1e9, nb_iter=8)
    assert np.all(np.isfinite(z0)) and np.all(np.isfinite(d0["J"]))
    # (the returned iterate is the extrapolated point, not the prox output -- SURVEY Q2 -- so it is
    # small but not zero under a huge lbda)
    assert np.all(np.isfinite(zh)) and np.max(np.abs(zh)) < np.max(np.abs(z0)) and np.all(np.isfinite(dh["theta"]))
