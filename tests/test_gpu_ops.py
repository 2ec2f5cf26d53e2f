"""GPU parity of the stand-alone operator kernels (through the C ABI) against the golden
vectors of the live reference and against the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pybold_oracle as orc  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


TOL = {np.float64: 1e-13, np.float32: 2e-6}


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_operators_match_reference_golden(golden, dt):
    import pybold_b200 as pb
    from pybold_b200 import convolution as cv
    g = golden("ops")
    for i in range(int(g["n_cases"])):
        k, x = g["k%d" % i].astype(dt), g["x%d" % i].astype(dt)
        T = len(x)
        tol = TOL[dt] * (10 if dt is np.float32 else 1)
        assert rel(cv.simple_convolve(k, x), g["conv%d" % i]) < tol
        assert rel(cv.spectral_convolve(k, x), g["sconv%d" % i]) < tol
        assert rel(cv.simple_retro_convolve(k, x), g["corr%d" % i]) < tol
        assert rel(cv.spectral_retro_convolve(k, x), g["scorr%d" % i]) < tol
        D = pb.DiscretInteg()
        assert rel(D.op(x), g["integ_op%d" % i]) < tol * 10
        assert rel(D.adj(x), g["integ_adj%d" % i]) < tol * 10
        H = pb.ConvAndLinear(pb.DiscretInteg(), k, dim_in=T, dim_out=T)
        assert rel(H.op(x), g["H_op%d" % i]) < tol * 10
        assert rel(H.adj(x), g["H_adj%d" % i]) < tol * 10
        out = cv.simple_convolve(k, x)
        assert isinstance(out, np.ndarray) and out.dtype == dt and out.shape == x.shape


def test_batched_operators_and_dot_test():
    """<A x, y> == <x, A^T y> on a ragged set of shapes, batched with per-voxel kernels."""
    import pybold_b200 as pb
    rng = np.random.RandomState(0)
    for (V, T, K) in [(1, 1, 1), (3, 7, 9), (65, 300, 20), (17, 1201, 28), (5, 33, 64)]:
        x = torch.as_tensor(rng.randn(V, T), device="cuda")
        yv = torch.as_tensor(rng.randn(V, T), device="cuda")
        k = torch.as_tensor(rng.randn(V, K), device="cuda")
        H = pb.ConvAndLinear(pb.DiscretInteg(), k, dim_in=T)
        lhs = torch.sum(H.op(x) * yv, dim=1)
        rhs = torch.sum(x * H.adj(yv), dim=1)
        assert torch.allclose(lhs, rhs, rtol=1e-11, atol=1e-9)
        for v in (0, V - 1):
            want = orc.conv_causal(k[v].cpu().numpy(), np.cumsum(x[v].cpu().numpy()))
            assert rel(H.op(x)[v].cpu().numpy(), want) < 1e-12


def test_spm_hrf_matches_reference_golden(golden):
    import pybold_b200 as pb
    g = golden("spm_hrf")
    for i, (delta, t_r, dur) in enumerate(g["grid"]):
        h, t = pb.spm_hrf(delta, t_r, dur, False)
        hn, _ = pb.spm_hrf(delta, t_r, dur, True)
        assert h.shape == g["h%d" % i].shape
        assert rel(h, g["h%d" % i]) < 1e-14
        assert rel(hn, g["hn%d" % i]) < 1e-14
        assert np.array_equal(t, g["t%d" % i])
    # batched thetas on the device
    th = torch.tensor([0.6, 1.0, 1.9], device="cuda", dtype=torch.float64)
    hb, _ = pb.spm_hrf(th, 1.0, 20.0, False)
    assert rel(hb[1].cpu().numpy(), g["h0"]) < 1e-14
    with pytest.raises(ValueError):
        pb.spm_hrf(2.5, 1.0, 20.0)
    with pytest.raises(ValueError):
        pb.spm_hrf(0.4, 1.0, 20.0)


def test_lipschitz_matches_reference_golden(golden):
    import pybold_b200 as pb
    from pybold_b200 import _lib
    from pybold_b200.utils import spectral_radius_est
    g = golden("lipschitz")
    for i, (T, _, _, _) in enumerate(g["cases"]):
        T = int(T)
        h = g["h%d" % i]
        H = pb.ConvAndLinear(pb.DiscretInteg(), h, dim_in=T)
        got = spectral_radius_est(H, (T,), x0=g["x0_%d" % i])
        assert abs(got / float(g["power%d" % i]) - 1) < 1e-12
        # global-RNG behaviour of the reference (utils.py:97)
        np.random.seed(100 + i)
        got2 = spectral_radius_est(H, (T,))
        assert abs(got2 / float(g["power%d" % i]) - 1) < 1e-12
        hd = torch.as_tensor(h, device="cuda")
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        rc = _lib.lib.pb_lipschitz_frob_f64(hd.data_ptr(), 0, out.data_ptr(), 1, T, len(h), 0)
        assert rc == 0
        assert abs(float(out[0]) / float(g["frob%d" % i]) - 1) < 1e-13


def test_frobenius_formula_edge_shapes():
    """Closed-form ||A^T A||_F against the dense Gram matrix, including K > T and K = 1."""
    from pybold_b200 import _lib
    rng = np.random.RandomState(5)
    for (T, K) in [(5, 1), (7, 2), (10, 12), (19, 20), (20, 20), (21, 20), (64, 33), (50, 64), (400, 27)]:
        h = rng.randn(K)
        hd = torch.as_tensor(h, device="cuda")
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        assert _lib.lib.pb_lipschitz_frob_f64(hd.data_ptr(), 0, out.data_ptr(), 1, T, K, 0) == 0
        assert abs(float(out[0]) / orc.frobenius_lipschitz(h, T) - 1) < 1e-12, (T, K)


def test_abi_error_codes():
    from pybold_b200 import _lib
    x = torch.zeros(8, device="cuda", dtype=torch.float64)
    assert _lib.lib.pb_integ_op_f64(0, x.data_ptr(), 1, 8, 0) == _lib.PB_ERR_INVALID_ARG
    assert _lib.lib.pb_integ_op_f64(x.data_ptr(), x.data_ptr(), 1, -3, 0) == _lib.PB_ERR_INVALID_ARG
    assert _lib.lib.pb_integ_op_f64(x.data_ptr(), x.data_ptr(), 1, 10 ** 6, 0) == _lib.PB_ERR_UNSUPPORTED
    assert _lib.lib.pb_integ_op_f64(x.data_ptr(), x.data_ptr(), 0, 8, 0) == 0      # empty batch is fine
    with pytest.raises(ValueError):
        _lib.check(_lib.PB_ERR_UNSUPPORTED, "x")


def test_layout_adapter_roundtrip_and_ragged_tiles():
    """[T, V] <-> [V, T] transpose kernel (row N4): exact, including partial 32 x 32 tiles."""
    from pybold_b200.io import timeseries_from_voxels, voxels_from_timeseries
    rng = np.random.RandomState(2)
    for (T, V, dt) in [(1, 1, np.float64), (33, 65, np.float32), (300, 1001, np.float32),
                       (1200, 4097, np.float64)]:
        a = rng.randn(T, V).astype(dt)
        vt = voxels_from_timeseries(a)
        assert vt.is_cuda and vt.shape == (V, T) and vt.is_contiguous()
        assert np.array_equal(vt.cpu().numpy(), a.T)
        back = timeseries_from_voxels(vt)
        assert np.array_equal(back.cpu().numpy(), a)
    with pytest.raises(ValueError):
        voxels_from_timeseries(np.zeros(5))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_layout_adapter_tma_tiles(dt):
    """The TMA tile mover (csrc/pb_transpose_tma.cuh) serves 16-byte aligned row pitches: full tiles,
    edge boxes clipped by the tensor maps, more tiles than CTAs x stages (mbarrier phases wrap), matrices
    smaller than one box, and a misaligned base pointer falling back to the plain kernel -- all exact."""
    from pybold_b200.io import timeseries_from_voxels, voxels_from_timeseries
    rng = np.random.RandomState(3)
    q = 4 if dt is np.float32 else 2
    for (T, V) in [(64, 64), (128, 256), (4 * q, 8 * q), (300, 1000), (1200, 4096), (68, 10 ** 5), (2048, 3000),
                   (36, 60)]:
        assert T % q == 0 and V % q == 0
        a = torch.as_tensor(rng.randn(T, V).astype(dt), device="cuda")
        vt = voxels_from_timeseries(a)
        assert vt.shape == (V, T) and torch.equal(vt, a.t().contiguous()), (T, V)
        assert torch.equal(timeseries_from_voxels(vt), a), (T, V)
    # same data at a base address that is only element-aligned: plain-load kernel, same answer
    buf = torch.as_tensor(rng.randn(300 * 1000 + 1).astype(dt), device="cuda")
    a = buf[1:].reshape(300, 1000)
    assert a.data_ptr() % 16 != 0
    from pybold_b200 import _lib
    out = torch.empty((1000, 300), dtype=a.dtype, device="cuda")
    assert _lib.fn("pb_transpose", a.dtype)(a.data_ptr(), out.data_ptr(), 300, 1000, 0) == 0
    torch.cuda.synchronize()
    assert torch.equal(out, a.t().contiguous())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_row_kernels_all_shapes(dt):
    """Register-resident row kernels (pb_ops_rows.cuh) and the generic fallback against NumPy float64:
    16-byte aligned rows (fast path), ragged rows and long kernels (fallback), shared and per-row
    taps, chunks partly or wholly beyond the row end."""
    import pybold_b200 as pb
    from pybold_b200 import convolution as cv
    rng = np.random.RandomState(11)
    tol = 3e-5 if dt is np.float32 else 1e-12
    shapes = [(9, 4, 1), (33, 128, 5), (70, 300, 20), (40, 384, 20), (21, 388, 28), (19, 600, 32),
              (13, 640, 28), (11, 1200, 28), (7, 1280, 20), (5, 1284, 20), (6, 301, 20), (4, 600, 40),
              (3, 2, 3)]
    for (V, T, K) in shapes:
        xn = rng.randn(V, T)
        for per_row in (False, True):
            kn = rng.randn(V, K) if per_row else rng.randn(K)
            x = torch.as_tensor(xn.astype(dt), device="cuda")
            k = torch.as_tensor(kn.astype(dt), device="cuda")
            x64 = x.cpu().numpy().astype(np.float64)
            k64 = np.broadcast_to(k.cpu().numpy().astype(np.float64), (V, K))
            conv = np.stack([orc.conv_causal(k64[v], x64[v]) for v in range(V)])
            corr = np.stack([np.correlate(np.concatenate([x64[v], np.zeros(K - 1)]), k64[v], "valid")
                             for v in range(V)])
            cs = np.cumsum(x64, axis=1)
            rcs = np.cumsum(x64[:, ::-1], axis=1)[:, ::-1]
            D = pb.DiscretInteg()
            H = pb.ConvAndLinear(D, k, dim_in=T)
            hop = np.stack([orc.conv_causal(k64[v], cs[v]) for v in range(V)])
            hadj = np.cumsum(corr[:, ::-1], axis=1)[:, ::-1]
            got = {"conv": cv.simple_convolve(k, x), "corr": cv.simple_retro_convolve(k, x),
                   "integ": D.op(x), "integ_adj": D.adj(x), "H.op": H.op(x), "H.adj": H.adj(x)}
            want = {"conv": conv, "corr": corr, "integ": cs, "integ_adj": rcs, "H.op": hop, "H.adj": hadj}
            for name in got:
                assert got[name].shape == (V, T)
                assert rel(got[name].cpu().numpy(), want[name]) < tol, (name, V, T, K, per_row, dt)


def test_row_kernels_in_place_and_misaligned():
    """The C ABI allows out == x; rows that do not start on 16-byte boundaries take the fallback."""
    from pybold_b200 import _lib
    rng = np.random.RandomState(12)
    V, T = 37, 300
    xn = rng.randn(V, T).astype(np.float32)
    want = np.cumsum(xn.astype(np.float64), axis=1)
    x = torch.as_tensor(xn, device="cuda")
    assert _lib.lib.pb_integ_op_f32(x.data_ptr(), x.data_ptr(), V, T, 0) == 0
    assert rel(x.cpu().numpy(), want) < 3e-5
    buf = torch.zeros(V * T + 1, device="cuda", dtype=torch.float32)
    buf[1:] = torch.as_tensor(xn, device="cuda").reshape(-1)
    out = torch.empty(V * T + 1, device="cuda", dtype=torch.float32)
    assert _lib.lib.pb_integ_op_f32(buf.data_ptr() + 4, out.data_ptr() + 4, V, T, 0) == 0
    assert rel(out[1:].reshape(V, T).cpu().numpy(), want) < 3e-5


def test_toeplitz_and_rectangular_convolutions_like_the_reference_tests():
    """pybold/tests/test_convolution.py:169-211: the Toeplitz matrix with the block signal as kernel
    (rectangular, dim_in = len(hrf)) reproduces the direct convolution and its adjoint."""
    import pybold_b200 as pb
    from pybold_b200 import convolution as cv
    rng = np.random.RandomState(21)
    for (T, K) in [(300, 20), (600, 30), (128, 7)]:
        ai_s = np.zeros(T)
        ai_s[rng.randint(0, T - 20, 4)] = 1.0
        ai_s = np.cumsum(ai_s) % 2.0                      # block signal
        hrf = pb.spm_hrf(1.0, 1.0, float(K), True)[0]
        assert len(hrf) == K
        H = cv.toeplitz_from_kernel(ai_s, len(hrf), len(ai_s))
        assert H.shape == (T, K)
        want = np.zeros((T, K))
        for i in range(T):
            for c in range(K):
                if 0 <= i - c < T:
                    want[i, c] = ai_s[i - c]
        assert np.array_equal(H, want)
        ar_ref = cv.simple_convolve(hrf, ai_s)
        assert np.allclose(ar_ref, H.dot(hrf), atol=1.0e-7)
        adj_ref = cv.simple_retro_convolve(ai_s, H.dot(hrf), len(hrf))
        assert adj_ref.shape == (K,)
        assert np.allclose(adj_ref, H.T.dot(H.dot(hrf)), atol=1.0e-7)
        # square case and dim_out > len(x)
        Hs = cv.toeplitz_from_kernel(hrf, T)
        assert np.allclose(Hs.dot(ai_s), cv.simple_convolve(hrf, ai_s), atol=1.0e-12)
        assert np.allclose(Hs.T.dot(ai_s), cv.simple_retro_convolve(hrf, ai_s), atol=1.0e-12)
        longer = cv.simple_convolve(hrf, ai_s[:50], 55)
        assert longer.shape == (55,) and np.allclose(longer, np.convolve(hrf, ai_s[:50])[:55], atol=1e-12)


def test_rectangular_conv_and_linear_matches_dense_toeplitz():
    """ConvAndLinear with dim_out != dim_in (pybold/linear.py:46-113) against K.dot(cumsum) / revcumsum(K.T.dot)."""
    import pybold_b200 as pb
    from pybold_b200 import convolution as cv
    rng = np.random.RandomState(22)
    for (dim_in, dim_out, K) in [(120, 100, 20), (100, 130, 12)]:
        k, x, yv = rng.randn(K), rng.randn(dim_in), rng.randn(dim_out)
        H = pb.ConvAndLinear(pb.DiscretInteg(), k, dim_in, dim_out)
        Kmat = cv.toeplitz_from_kernel(k, dim_in, dim_out)
        assert np.allclose(H.op(x), Kmat.dot(np.cumsum(x)), atol=1e-11)
        assert np.allclose(H.adj(yv), np.cumsum(Kmat.T.dot(yv)[::-1])[::-1], atol=1e-11)
        assert abs(np.dot(H.op(x), yv) - np.dot(x, H.adj(yv))) < 1e-9


def test_spm_hrf_shape_parameters_vs_reference_golden(golden):
    """Non-default delay / dispersion / ratio / onset / dt (pybold/hrf_model.py:12-14) on `pb_spm_hrf_ex_*`."""
    import pybold_b200 as pb
    g = golden("spm_hrf_params")
    for i in range(int(g["n"])):
        kw = {k: float(v) for k, v in g["kw%d" % i]}
        for norm, key in ((False, "h"), (True, "hn")):
            h, t = pb.spm_hrf(normalized_hrf=norm, **kw)
            want = g["%s%d" % (key, i)]
            assert h.shape == want.shape and np.array_equal(t, g["t%d" % i])
            assert np.max(np.abs(h - want)) <= 1e-12 * (np.max(np.abs(want)) + 1e-300), (i, norm)
    # batched theta, and the default parameters spelled out take the fast kernel with the same answer
    th = np.array([0.6, 1.0, 1.9])
    hb, _ = pb.spm_hrf(th, t_r=0.75, dur=20.0, normalized_hrf=False, p_delay=5.5)
    for v, t in enumerate(th):
        h1, _ = pb.spm_hrf(float(t), t_r=0.75, dur=20.0, normalized_hrf=False, p_delay=5.5)
        assert np.array_equal(hb[v], h1)
    with pytest.raises(ValueError):
        pb.spm_hrf(1.0, p_disp=0.0)


def test_spectral_radius_est_on_any_operator():
    """utils.py:94-109 takes any object with op / adj: the integration operator alone (known largest singular
    value of the T x T summation matrix) and a user-defined operator give the reference's numbers."""
    import pybold_b200 as pb
    from pybold_b200.utils import spectral_radius_est
    T = 64
    x0 = np.random.RandomState(4).randn(T)
    got = spectral_radius_est(pb.DiscretInteg(), (T,), x0=x0, nb_iter=60)
    Lmat = np.tril(np.ones((T, T)))

    class Dense:
        def op(self, x):
            return Lmat.dot(x)

        def adj(self, x):
            return Lmat.T.dot(x)
    want = orc.spectral_radius_est(Dense(), x0, nb_iter=60)
    assert abs(got / want - 1) < 1e-12
    assert abs(spectral_radius_est(Dense(), (T,), x0=x0, nb_iter=60) / want - 1) < 1e-14
    np.random.seed(5)
    a = spectral_radius_est(pb.DiscretInteg(), (T,))
    np.random.seed(5)
    assert abs(a / orc.spectral_radius_est(Dense(), np.random.randn(T)) - 1) < 1e-12


def test_layout_adapter_streams_host_matrices():
    """Host [T, V] input in column chunks (pinned staging, two streams): same result as the one-shot path,
    including a ragged last chunk, float64, a CPU tensor."""
    from pybold_b200.io import voxels_from_timeseries
    rng = np.random.RandomState(6)
    a = rng.randn(300, 1000).astype(np.float32)
    got = voxels_from_timeseries(a, chunk_voxels=256)            # 3 full chunks + 232 voxels
    assert got.is_cuda and got.shape == (1000, 300) and np.array_equal(got.cpu().numpy(), a.T)
    b = rng.randn(64, 513)
    assert np.array_equal(voxels_from_timeseries(torch.from_numpy(b), chunk_voxels=128).cpu().numpy(), b.T)
    assert np.array_equal(voxels_from_timeseries(a, chunk_voxels=4096).cpu().numpy(), a.T)    # one chunk: one-shot path
