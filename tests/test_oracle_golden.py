"""Pin the CPU oracle (oracle/pybold_oracle.py) against golden vectors produced by the live
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pybold_oracle as orc


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


def test_linear_operators(golden):
    g = golden("ops")
    for i in range(int(g["n_cases"])):
        k, x = g["k%d" % i], g["x%d" % i]
        # reference tests use atol=1e-7 (pybold/tests/test_convolution.py:29-36); we hold 1e-13
        for ref_key in ("conv", "sconv", "toep_dot"):
            assert rel(orc.conv_causal(k, x), g[ref_key + str(i)]) < 1e-13
        for ref_key in ("corr", "scorr", "toepT_dot"):
            assert rel(orc.corr_anticausal(k, x), g[ref_key + str(i)]) < 1e-13
        assert np.array_equal(orc.integ_op(x), g["integ_op%d" % i])
        assert np.array_equal(orc.integ_adj(x), g["integ_adj%d" % i])
        H = orc.HrfIntegOperator(k, len(x))
        assert rel(H.op(x), g["H_op%d" % i]) < 1e-14
        assert rel(H.adj(x), g["H_adj%d" % i]) < 1e-14
        assert rel(orc.toeplitz_from_kernel(k, len(x)).dot(x), g["toep_dot%d" % i]) == 0.0


def test_spm_hrf(golden):
    g = golden("spm_hrf")
    for i, (delta, t_r, dur) in enumerate(g["grid"]):
        h, t = orc.spm_hrf(delta, t_r, dur, False)
        hn, _ = orc.spm_hrf(delta, t_r, dur, True)
        assert np.array_equal(h, g["h%d" % i])
        assert np.array_equal(hn, g["hn%d" % i])
        assert np.array_equal(t, g["t%d" % i])
        assert len(h) == orc.hrf_len(t_r, dur)
        # closed form at the kept samples (what the device evaluates)
        assert rel(orc.spm_hrf_closed_form(delta, t_r, dur), g["h%d" % i]) < 5e-15
        h_cf, _, _ = orc.hrf_taps_and_derivs(delta, t_r, dur)
        assert rel(h_cf, g["h%d" % i]) < 5e-15


def test_spm_hrf_rejects_out_of_range():
    for bad in (0.49, 2.01):
        with pytest.raises(ValueError):
            orc.spm_hrf(bad, 1.0, 20.0)


def test_hrf_derivatives_match_finite_differences():
    for theta in (0.65, 1.0, 1.7):
        h, h1, h2 = orc.hrf_taps_and_derivs(theta, 0.75, 20.0)
        e = 1e-5
        hp, h1p, _ = orc.hrf_taps_and_derivs(theta + e, 0.75, 20.0)
        hm, h1m, _ = orc.hrf_taps_and_derivs(theta - e, 0.75, 20.0)
        assert rel((hp - hm) / (2 * e), h1) < 1e-8
        assert rel((h1p - h1m) / (2 * e), h2) < 1e-8


def test_lipschitz(golden):
    g = golden("lipschitz")
    for i, (T, _, _, _) in enumerate(g["cases"]):
        T = int(T)
        h = g["h%d" % i]
        H = orc.HrfIntegOperator(h, T)
        assert orc.spectral_radius_est(H, g["x0_%d" % i]) == float(g["power%d" % i])
        assert abs(orc.frobenius_lipschitz(h, T) / float(g["frob%d" % i]) - 1) < 1e-15


def test_deconv_fixed_lambda(golden):
    g = golden("deconv_fixed")
    for tag in g["tags"]:
        y = g["y"][int(g["voxel_" + tag])]
        x, z, w, J, n_done = orc.deconv_fixed_lbda(
            y, g["h"], float(g["lbda_" + tag]), x0_power=g["x0_" + tag],
            early_stopping=bool(g["es_" + tag]), tol=float(g["tol_" + tag]),
            wind=int(g["wind_" + tag]), nb_iter=int(g["nb_iter_" + tag]))
        assert n_done == len(g["J_" + tag]), tag          # same early-stop iteration (Q5)
        assert np.array_equal(w, g["dz_" + tag]), tag     # bit exact
        assert np.array_equal(z, g["z_" + tag]), tag
        assert rel(x, g["x_" + tag]) < 1e-14, tag          # reference x goes through the FFT
        assert rel(J, g["J_" + tag]) < 1e-13, tag
    assert len(g["J_c"]) < 400 and len(g["J_d"]) < 1000    # the early stops did fire


def test_loops_deconv(golden):
    g = golden("loops_deconv")
    for tag in g["tags"]:
        v, _, lbda, n, es, tol = g["par_" + tag]
        w = orc.loops_deconv(g["y"][int(v)], g["w0_" + tag], g["h_" + tag], lbda, int(n),
                             bool(es), tol)
        # Numba's BLAS summation order differs from NumPy's by rounding only (survey: 3e-17)
        assert np.max(np.abs(w - g["w_" + tag])) < 1e-15, tag


def test_hrf_fit_err(golden):
    g = golden("hrf_fit_err")
    for th, val in zip(g["thetas"], g["vals"]):
        got = orc.hrf_fit_err(th, g["z"], g["y"], float(g["t_r"]), float(g["dur"]))
        assert abs(got / val - 1) < 1e-13
        got2 = orc.hrf_fit_err_fast(th, g["z"], g["y"], float(g["t_r"]), float(g["dur"]))
        assert abs(got2 / val - 1) < 1e-13


def _bd_kwargs(g, tag):
    kw = {}
    for key in ("lbda", "hrf_dur", "nb_iter", "theta_0", "early_stopping", "wind", "tol"):
        name = key + "_" + tag
        if name in g.files:
            val = g[name]
            kw[key] = val.item()
    if "z_0_" + tag in g.files:
        kw["z_0"] = g["z_0_" + tag]
    return kw


@pytest.mark.parametrize("tag", ["t300_v1", "t240_warm", "t1200_v0", "t300_flat", "t300_es"])
def test_bd_with_reference_theta_solver(golden, tag):
    """Same SciPy L-BFGS-B call as the reference => the whole trajectory reproduces."""
    g = golden("bd")
    trace = {}
    x, z, w, h, d = orc.bd(g["y_" + tag], float(g["t_r_" + tag]), theta_solver="lbfgsb",
                           trace=trace, **_bd_kwargs(g, tag))
    assert len(d["J"]) == len(g["J_" + tag])
    # the reference's own conditioning is ~3e-8 for a 1e-13 perturbation
    # (measured: theta 2.4e-8, z/h 9e-9, J 1.4e-9 at T <= 300; 3.9e-7, 2.8e-7, 5.6e-9 at T = 1200)
    long = tag.startswith("t1200")
    assert np.max(np.abs(np.array(trace["theta"]) - g["thetas_" + tag])) < (4e-6 if long else 2.5e-7)
    assert rel(z, g["z_" + tag]) < (3e-6 if long else 1e-7)
    assert rel(h, g["h_" + tag]) < (3e-6 if long else 1e-7)
    assert rel(d["J"], g["J_" + tag]) < (6e-8 if long else 1.5e-8)
    assert rel(d["r"], g["r_" + tag]) < (6e-8 if long else 1.5e-8)
    assert rel(d["g"], g["g_" + tag]) < (2.5e-6 if long else 1e-7)


def test_theta_step_exact_vs_reference_lbfgsb(golden):
    """The reference's theta is within ~1e-7 of the exact bounded minimiser."""
    g = golden("bd")
    for tag in ("t300_v0", "t240_v0", "t1200_v0"):
        t_r = float(g["t_r_" + tag])
        dur = float(g["hrf_dur_" + tag])
        thetas = g["thetas_" + tag]
        zs = g["zs_" + tag]
        prev = 2.0
        worst = 0.0
        for i in range(0, len(thetas), max(1, len(thetas) // 12)):
            prev = thetas[i - 1] if i else 2.0
            th = orc.theta_step_exact(prev, zs[i], g["y_" + tag], t_r, dur, [(0.6, 1.9)])
            worst = max(worst, abs(th - thetas[i]))
            # and it is a stationary point / bound of the exact cost
            f0 = orc.hrf_fit_err_fast(th, zs[i], g["y_" + tag], t_r, dur)
            for e in (-1e-5, 1e-5):
                if 0.6 <= th + e <= 1.9:
                    assert orc.hrf_fit_err_fast(th + e, zs[i], g["y_" + tag], t_r, dur) >= f0
        assert worst < 5e-7, (tag, worst)


@pytest.mark.parametrize("tag", ["t300_v0", "t240_v0"])
def test_bd_exact_theta_end_to_end(golden, tag):
    """bd with the exact theta step (device algorithm) against the reference, stated tolerance."""
    g = golden("bd")
    trace = {}
    x, z, w, h, d = orc.bd(g["y_" + tag], float(g["t_r_" + tag]), theta_solver="exact",
                           trace=trace, **_bd_kwargs(g, tag))
    # measured: theta 5.3e-8, z 6.0e-8, h 6.1e-8, J 3.9e-9
    assert np.max(np.abs(np.array(trace["theta"]) - g["thetas_" + tag])) < 5e-7
    assert rel(z, g["z_" + tag]) < 6e-7
    assert rel(h, g["h_" + tag]) < 6e-7
    assert rel(d["J"], g["J_" + tag]) < 4e-8


# ---- rows A8 / N1 / N2: vectors of the live reference added in round 2 ---------------------------
def _auto_kwargs(g, tag):
    return dict(early_stopping=bool(g["early_stopping_" + tag]), tol=float(g["tol_" + tag]),
                wind=int(g["wind_" + tag]), nb_iter=int(g["nb_iter_" + tag]),
                nb_sub_iter=int(g["nb_sub_iter_" + tag]))


def test_deconv_auto_lambda_vs_reference(golden):
    """``deconv(lbda=None)`` of the live reference with sigma injected: the oracle's
    restatement of bold_signal.py:99-214 reproduces it to rounding, outer stop included."""
    g = golden("deconv_auto")
    stopped = 0
    for tag in g["tags"]:
        kw = _auto_kwargs(g, tag)
        x, z, w, J, R, G, _ = orc.deconv_auto_lbda(g["y_" + tag], g["h_" + tag], float(g["sigma_" + tag]),
                                                   x0_power=g["x0_" + tag], **kw)
        assert len(J) == len(g["J_" + tag]), tag                  # same alpha-window stop iteration
        stopped += len(J) < kw["nb_iter"]
        assert np.max(np.abs(w - g["dz_" + tag])) < 1e-18, tag
        assert np.array_equal(z, g["z_" + tag]), tag
        assert rel(x, g["x_" + tag]) < 1e-14, tag                  # reference x goes through the FFT
        for got, key in ((J, "J_"), (R, "R_"), (G, "G_")):
            assert rel(got, g[key + tag]) < 1e-14, (tag, key)
    assert stopped >= 2                                            # the alpha window did fire


def test_hrf_estim_vs_reference(golden):
    g = golden("hrf_estim")
    for tag in g["tags"]:
        args = (g["z_" + tag], g["y_" + tag], float(g["t_r_" + tag]), float(g["dur_" + tag]))
        h, J, theta = orc.hrf_estim(*args)                         # same SciPy call as the reference
        assert rel(h, g["h_" + tag]) < 1e-7 and len(J) == len(g["J_" + tag]), tag
        assert rel(J, g["J_" + tag]) < 1e-9, tag
        he, Je, the = orc.hrf_estim_exact(*args)                   # the device algorithm
        assert abs(the - theta) < 5e-7, tag                        # measured <= 8e-8
        assert rel(he, g["h_" + tag]) < 2e-6, tag                  # measured <= 2.2e-7
        assert abs(Je[-1] / g["J_" + tag][-1] - 1) < 1e-12, tag    # flat at the minimum


def test_spm_hrf_shape_parameters(golden):
    g = golden("spm_hrf_params")
    for i in range(int(g["n"])):
        kw = {k: float(v) for k, v in g["kw%d" % i]}
        for norm, key in ((False, "h"), (True, "hn")):
            h, t = orc.spm_hrf(normalized_hrf=norm, **kw)
            assert np.array_equal(h, g["%s%d" % (key, i)]) and np.array_equal(t, g["t%d" % i])
            hc, tc = orc.spm_hrf_general(normalized_hrf=norm, **kw)   # what the device evaluates
            scale = np.max(np.abs(h)) + 1e-300
            assert np.max(np.abs(hc - h)) / scale < 1e-12, (i, norm)
            assert np.max(np.abs(tc - t)) < 1e-9


def test_bd_t1200_full_iterations_exact_theta(golden):
    """cfg4 shape with the reference's nb_iter = 100: theta walks to the lower bound 0.6 and stays."""
    g = golden("bd_t1200")
    tag = "t1200_n100"
    trace = {}
    x, z, w, h, d = orc.bd(g["y_" + tag], float(g["t_r_" + tag]), theta_solver="exact", trace=trace,
                           **_bd_kwargs(g, tag))
    assert np.max(np.abs(np.array(trace["theta"]) - g["thetas_" + tag])) < 1e-5    # measured 1.0e-6
    assert g["thetas_" + tag][-1] == 0.6 and trace["theta"][-1] == 0.6
    assert rel(z, g["z_" + tag]) < 3e-7 and rel(x, g["x_" + tag]) < 3e-7           # measured 2.9e-8
    assert rel(w, g["dz_" + tag]) < 2e-6 and rel(h, g["h_" + tag]) < 1e-12
    assert rel(d["J"], g["J_" + tag]) < 4e-7 and rel(d["g"], g["g_" + tag]) < 2.5e-6


def test_dwt_convention_on_pywavelets_documented_examples():
    """PyWavelets is absent, so the DWT convention of ``mad_daub_noise_est`` (utils.py:22) is pinned on
    the Haar examples of PyWavelets' documentation: ``pywt.dwt([1, 2, 3, 4, 5, 6], 'db1')`` gives
    ``cD = [-0.70710678] * 3``, ``pywt.wavedec(range(1, 9), 'db1', level=2)`` has
    ``cD1 = [-0.70710678] * 4``; the default mode mirrors ``... x2 x1 | x1 x2 ... xn | xn xn-1 ...`` and
    a mode-'symmetric' transform of N samples with an F-tap filter has floor((N + F - 1) / 2)
    coefficients.  (Quoted from the documentation, not generated here.)"""
    r = np.sqrt(0.5)
    assert np.allclose(orc.dwt_detail_level1([1, 2, 3, 4, 5, 6], orc._DB1_DEC_HI), [-r] * 3, atol=1e-15)
    assert np.allclose(orc.dwt_detail_level1(np.arange(1, 9), orc._DB1_DEC_HI), [-r] * 4, atol=1e-15)
    assert np.allclose(orc.dwt_detail_level1([1, 2, 3], orc._DB1_DEC_HI), [-r, 0.0], atol=1e-15)
    for T in (10, 11, 300, 301):
        assert len(orc.db3_detail_level1(np.arange(T, dtype=float))) == (T + 5) // 2
    # three vanishing moments: a quadratic is annihilated away from the borders
    t = np.arange(64, dtype=float)
    cD = orc.db3_detail_level1(3.0 + 0.5 * t - 0.01 * t * t)
    assert np.max(np.abs(cD[3:-3])) < 1e-10 and np.max(np.abs(cD[:3])) > 1e-3
    # short series: the reference's level-0 fallback (utils.py:23-24) is the MAD of the series itself
    x = np.array([0.3, -1.0, 2.0, 0.1, 0.7])
    assert orc.mad_daub_noise_est(x) == orc.mad(x)
