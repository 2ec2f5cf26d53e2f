"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, the product never touches the oracle, and the host-side mirror keeps the
reference's signatures."""
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "pybold_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    from pybold_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(_lib.lib, name), "libpybold_b200.so does not export " + name
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    assert _lib.lib.pb_version() >= 100
    assert _lib.lib.pb_max_T() >= 1200 and _lib.lib.pb_max_K() >= 32


def test_hrf_len_matches_reference_rule():
    from pybold_b200 import _lib
    for t_r, dur, K in [(1.0, 20.0, 20), (0.75, 20.0, 27), (0.72, 20.0, 28), (0.7535, 20.0, 27),
                        (1.0, 30.0, 30), (2.0, 60.0, 30), (0.1, 10.0, 100)]:
        assert _lib.lib.pb_hrf_len(t_r, dur) == K
        assert K == len(range(0, int(dur / 0.001), int(t_r / 0.001)))
    assert _lib.lib.pb_hrf_len(0.0, 20.0) == _lib.PB_ERR_INVALID_ARG


def test_error_strings_and_exceptions():
    from pybold_b200 import _lib
    assert _lib.error_string(0) == "ok"
    assert "invalid" in _lib.error_string(_lib.PB_ERR_INVALID_ARG)
    with pytest.raises(ValueError):
        _lib.check(_lib.PB_ERR_INVALID_ARG, "x")
    with pytest.raises(_lib.PyboldB200Error):
        _lib.check(700, "x")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pybold_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "pybold_oracle" not in text, f


def test_signatures_mirror_the_reference():
    import pybold_b200 as pb
    want_bd = ["y", "t_r", "lbda", "theta_0", "z_0", "hrf_dur", "bounds", "nb_iter", "nb_sub_iter",
               "nb_last_iter", "print_period", "early_stopping", "wind", "tol", "verbose"]
    got = list(inspect.signature(pb.bd).parameters)
    assert got[:len(want_bd)] == want_bd
    sig = inspect.signature(pb.bd).parameters
    assert sig["lbda"].default == 1.0 and sig["nb_iter"].default == 100 and sig["wind"].default == 4
    assert sig["tol"].default == 1.0e-12 and sig["early_stopping"].default is False
    want_dc = ["y", "t_r", "hrf", "lbda", "early_stopping", "tol", "wind", "nb_iter", "nb_sub_iter",
               "verbose"]
    got = list(inspect.signature(pb.deconv).parameters)
    assert got[:len(want_dc)] == want_dc
    sig = inspect.signature(pb.deconv).parameters
    assert sig["lbda"].default is None and sig["wind"].default == 6 and sig["nb_iter"].default == 1000
    want_hrf = ["delta", "t_r", "dur", "normalized_hrf", "dt", "p_delay", "undershoot", "p_disp",
                "u_disp", "p_u_ratio", "onset"]
    assert list(inspect.signature(pb.spm_hrf).parameters) == want_hrf
    assert list(inspect.signature(pb.ConvAndLinear.__init__).parameters)[1:] == \
        ["M", "kernel", "dim_in", "dim_out", "spectral_conv"]
    assert pb.MIN_DELTA == 0.5 and pb.MAX_DELTA == 2.0


def test_no_cpu_fallback_without_device():
    import torch
    import pybold_b200 as pb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        pb.bd(np.zeros(16), 1.0)


def test_synth_generator_is_rank_independent():
    from pybold_b200.synth import gen_voxels
    a = gen_voxels(6, 50, 1.0, 20.0, seed0=3)
    b = gen_voxels(3, 50, 1.0, 20.0, seed0=3, first_voxel=3)
    assert np.array_equal(a[3:], b)


def test_dispatch_covers_every_series_length():
    """Host-side dispatch (no GPU needed): every T <= 2560 with K <= 28 has a register-tiled FP32 variant
    whose slots hold the series, and the tuned shapes pick the kernels DESIGN.md names."""
    from pybold_b200 import _lib
    for K in (20, 27, 28):
        for T in list(range(1, 400)) + list(range(400, 2561, 7)) + [2560]:
            vid = _lib.lib.pb_solver_variant(T, K, 0)
            assert vid != 0, (T, K)
            lanes, R, kmax = vid // 1000000, (vid // 1000) % 1000, vid % 1000
            assert lanes * R >= T and kmax >= K, (T, K, vid)
            assert lanes * R < 2 * T + 2 * R + 320, (T, K, vid)      # never a grossly oversized variant
    assert _lib.lib.pb_solver_variant(300, 20, 0) == 16019020        # cfg3: two voxels per warp
    assert _lib.lib.pb_solver_variant(240, 27, 0) == 16015028        # ICASSP native shape
    assert _lib.lib.pb_solver_variant(1200, 28, 0) == 64020028       # cfg4: two warps per voxel
    assert _lib.lib.pb_solver_variant(600, 20, 0) == 32020020
    assert _lib.lib.pb_solver_variant(150, 20, 0) // 1000000 == 8    # four voxels per warp
    assert _lib.lib.pb_solver_variant(2000, 28, 0) == 128016028     # four warps, R = 16 (skewed layout)
    assert _lib.lib.pb_solver_variant(3000, 20, 0) == 192016020      # six warps per voxel
    # round 2: every K <= 64 (TR down to 0.32 s at hrf_dur = 20 s) and T <= 4096 has a register-tiled variant
    for K in (29, 40, 41, 64):
        for T in list(range(1, 200, 3)) + list(range(200, 4097, 37)) + [4096]:
            vid = _lib.lib.pb_solver_variant(T, K, 0)
            assert vid != 0, (T, K)
            lanes, R, kmax = vid // 1000000, (vid // 1000) % 1000, vid % 1000
            assert lanes * R >= T and kmax >= K, (T, K, vid)
    for K in (20, 28):
        for T in range(2561, 4097, 41):
            assert _lib.lib.pb_solver_variant(T, K, 0) // 1000000 in (192, 256), (T, K)
    assert _lib.lib.pb_solver_variant(300, 40, 0) == 32012040
    assert _lib.lib.pb_solver_variant(300, 65, 0) == 0               # beyond PB_MAX_K: generic kernel


def test_db3_highpass_filter_has_its_defining_properties():
    """The db3 decomposition high-pass used by `mad_daub_noise_est` (pybold/utils.py:16-25) cannot be pinned
    against PyWavelets here (not installed); its coefficients are pinned by what defines the filter: unit
    energy, three vanishing moments, orthogonality to its even shifts, and the quadrature-mirror relation to a
    low-pass that sums to sqrt(2).  (The boundary convention of pywt's symmetric mode stays unverified; the
    MAD of the detail coefficients is insensitive to it.)"""
    from pybold_b200.noise import _DB3_DEC_HI
    g = np.array(_DB3_DEC_HI)
    k = np.arange(len(g))
    assert abs(np.sum(g * g) - 1.0) < 1e-10                    # (tabulated to ~12 digits)
    for m in range(3):
        assert abs(np.sum(k ** m * g)) < 1e-10, m
    assert abs(np.sum(k ** 3 * g)) > 1e-2                      # exactly three vanishing moments
    for shift in (2, 4):
        assert abs(np.sum(g[shift:] * g[:-shift])) < 1e-10
    lo = g[::-1] * (-1.0) ** k                                 # quadrature mirror
    assert abs(abs(np.sum(lo)) - np.sqrt(2.0)) < 1e-10
    for shift in (0, 2, 4):
        hi_s = np.concatenate([np.zeros(shift), g])[:len(g)]
        assert abs(np.sum(lo * hi_s)) < 1e-10


def test_host_side_parameter_plumbing():
    """Host logic that needs no GPU: the Lipschitz constant is 0.9 x the estimate with the product taken in double
    (one expression for every deconv path), scalar parameters stay host floats until the launch, arrays are
    checked against the batch size, NumPy uploads keep values and dtype (CPU device = the plain path)."""
    import numpy as np
    import torch
    from pybold_b200._array import per_voxel, upload
    from pybold_b200.bold_signal import _lipschitz_cst
    est32 = float(np.float32(1234.5678))
    c = _lipschitz_cst(est32, torch.float32)
    assert isinstance(c, float) and c == 0.9 * est32
    cb = _lipschitz_cst(np.array([1.0, 2.0]), torch.float64)
    assert np.allclose(cb, [0.9, 1.8])
    ct = _lipschitz_cst(torch.tensor([3.0], dtype=torch.float32), torch.float32)
    assert ct.dtype == torch.float32 and float(ct[0]) == float(np.float32(0.9 * 3.0))
    t = upload(np.arange(6, dtype=np.float64).reshape(2, 3), torch.float32, "cpu")
    assert t.dtype == torch.float32 and t.shape == (2, 3) and t.is_contiguous() and float(t[1, 2]) == 5.0
    v, stride = per_voxel(np.linspace(0.5, 1.5, 4), 4, torch.float64, "cpu", "lbda")
    assert stride == 1 and v.shape == (4,) and float(v[3]) == 1.5
    s, stride = per_voxel(1.7, 4, torch.float64, "cpu", "lbda")
    assert stride == 0 and s.numel() == 1 and float(s[0]) == 1.7
    assert per_voxel(1.7, 9, torch.float64, "cpu", "lbda")[0] is s        # cached: no upload per call
    with pytest.raises(ValueError):
        per_voxel(np.ones(3), 4, torch.float64, "cpu", "lbda")
