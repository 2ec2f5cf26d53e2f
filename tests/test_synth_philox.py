"""Row N3: the "philox-v1" synthetic generator -- CPU checks of its NumPy statement (known-answer
vectors of Philox4x32-10, batch-split independence, SNR rule) and GPU parity of the CUDA kernel,
the inf-norm normalisation and the relative-error metric."""
import numpy as np
import pytest

from pybold_b200.synth import gen_voxels_philox, philox4x32


def test_philox4x32_known_answers():
    """Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in philox4x32(ctr, key)) == want
    # vectorised over counters
    out = philox4x32((np.arange(3), 0, 0, 0), (0, 0))
    assert int(out[0][0]) == 0x6627e8d5 and out[0].shape == (3,)


def test_philox_generator_properties():
    y, z, delta = gen_voxels_philox(64, 300, 1.0, 20.0, snr_db=10.0, seed=5)
    assert y.shape == (64, 300) and z.shape == (64, 300) and delta.shape == (64,)
    assert np.all((delta >= 0.7) & (delta <= 1.3))
    assert np.all(z == np.round(z)) and z.max() <= 5 and np.all(z.sum(axis=1) == 5 * 12)
    # any voxel range reproduces the same numbers
    y2, z2, d2 = gen_voxels_philox(20, 300, 1.0, 20.0, snr_db=10.0, seed=5, first_voxel=30)
    assert np.array_equal(y[30:50], y2) and np.array_equal(z[30:50], z2) and np.array_equal(delta[30:50], d2)
    # a different seed gives a different batch
    y3, _, _ = gen_voxels_philox(4, 300, 1.0, 20.0, seed=6)
    assert not np.allclose(y3, y[:4])
    # SNR rule of pybold/data.py:439-444: ||x|| / ||n|| = 10^(snr/20)
    yc, _, _ = gen_voxels_philox(8, 240, 0.75, 20.0, snr_db=200.0, seed=5)      # practically clean
    yn, _, _ = gen_voxels_philox(8, 240, 0.75, 20.0, snr_db=6.0, seed=5)
    ratio = np.linalg.norm(yc, axis=1) / np.linalg.norm(yn - yc, axis=1)
    assert np.allclose(ratio, 10.0 ** (6.0 / 20.0), rtol=1e-6)
    # the noise is standard normal before scaling
    _, _, _ = gen_voxels_philox(1, 8, 1.0, 20.0, seed=1)
    big, _, _ = gen_voxels_philox(4, 4000, 1.0, 20.0, snr_db=0.0, nb_events=0, seed=9)
    assert np.allclose(big, 0.0)              # no events: x = 0, so the scaled noise vanishes too


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_device_generator_matches_numpy_statement(dt):
    import torch
    from pybold_b200.synth import gen_voxels_device
    tdt = getattr(torch, dt)
    tol = 1e-11 if dt == "float64" else 2e-6
    for (V, T, t_r, ne, first) in [(37, 300, 1.0, 5, 0), (9, 240, 0.75, 5, 100000), (5, 1201, 0.72, 7, 2 ** 33),
                                   (3, 17, 1.0, 1, 7), (4, 64, 2.0, 0, 1)]:
        y, z, d = gen_voxels_device(V, T, t_r, 20.0, snr_db=7.0, nb_events=ne, seed=1234567891011,
                                    first_voxel=first, dtype=tdt, return_truth=True)
        yo, zo, do = gen_voxels_philox(V, T, t_r, 20.0, snr_db=7.0, nb_events=ne, seed=1234567891011,
                                       first_voxel=first)
        assert y.is_cuda and y.dtype == tdt and y.shape == (V, T)
        assert np.array_equal(z.cpu().numpy(), zo)                     # integer block signal: exact
        assert np.max(np.abs(d.cpu().numpy() - do)) < tol
        scale = max(np.max(np.abs(yo)), 1.0)
        assert np.max(np.abs(y.cpu().numpy() - yo)) / scale < tol, (V, T, t_r, ne)
    y_only = gen_voxels_device(6, 300, dtype=tdt)
    assert y_only.shape == (6, 300)
    with pytest.raises(ValueError):
        gen_voxels_device(2, 300, nb_events=40)


@pytest.mark.gpu
def test_device_generator_feeds_the_solver():
    """The simulation loop of examples/icassp_2019/simulation.py on the device: generate, solve,
    normalise, score -- nothing touches the host until the error table is read."""
    import torch
    import pybold_b200 as pb
    from pybold_b200.synth import gen_voxels_device
    from pybold_b200.utils import inf_norm, rel_l2_err
    y, z, _ = gen_voxels_device(64, 240, 0.75, 20.0, snr_db=15.0, seed=3, delta_range=(0.7, 0.7),
                                dtype=torch.float64, return_truth=True)
    _, est_z, _, est_h, _ = pb.bd(y, 0.75, lbda=1.6, theta_0=2.0, hrf_dur=20.0, bounds=[(0.6, 1.9)],
                                  nb_iter=30)
    nz, nh = inf_norm([est_z, est_h])
    assert nz.is_cuda and float(nz.abs().max(dim=1).values.min()) > 0.999999
    h_true, _ = pb.spm_hrf(0.7, 0.75, 20.0, True)
    err_h = rel_l2_err(nh, torch.as_tensor(h_true, device="cuda"))
    err_z = rel_l2_err(nz, inf_norm(z))
    assert err_h.shape == (64,) and err_z.shape == (64,)
    assert bool(torch.isfinite(err_h).all()) and bool(torch.isfinite(err_z).all())
    assert float(err_h.max()) < 1.0          # 30 outer iterations already move h towards the true HRF
    # against NumPy
    want = np.linalg.norm(nh.cpu().numpy() - h_true[None, :], axis=1) / np.linalg.norm(h_true)
    assert np.allclose(err_h.cpu().numpy(), want, rtol=1e-12)


@pytest.mark.gpu
def test_inf_norm_matches_reference_rule():
    import torch
    from pybold_b200.utils import inf_norm
    rng = np.random.RandomState(4)
    for dt, tol in ((np.float64, 1e-15), (np.float32, 1e-6)):
        a = rng.randn(13, 301).astype(dt)
        want = a / (np.max(np.abs(a), axis=1, keepdims=True) + 1.0e-12)
        got = inf_norm(a)
        assert isinstance(got, np.ndarray) and got.dtype == dt
        assert np.max(np.abs(got - want)) < tol
        v = rng.randn(77).astype(dt)
        assert np.max(np.abs(inf_norm(v) - v / (np.max(np.abs(v)) + 1.0e-12))) < tol
        cols = inf_norm(a, axis=0)
        assert np.max(np.abs(cols - a / (np.max(np.abs(a), axis=0, keepdims=True) + 1.0e-12))) < tol
    cube = rng.randn(3, 4, 5)
    assert np.max(np.abs(inf_norm(cube) - cube / (np.max(np.abs(cube)) + 1.0e-12))) < 1e-15
    with pytest.raises(ValueError):
        inf_norm(np.zeros((2, 2, 2, 2)))
    zero = inf_norm(torch.zeros(3, 8, device="cuda"))
    assert float(zero.abs().max()) == 0.0
