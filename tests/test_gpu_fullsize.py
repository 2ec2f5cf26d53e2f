"""Full-size runs (BASELINE.json shapes) checked through size-independent properties: the oracle
cannot run 100 000 voxels, but (a) a voxel's answer must not depend on the batch it is solved in
or on its position in it, (b) the operator pair must stay adjoint, (c) a sample of voxels taken
out of the big batch must match the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pybold_oracle as orc  # noqa: E402
from pybold_b200.synth import gen_voxels_chunked  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


def test_cfg3_full_batch_permutation_and_batch_independence():
    """cfg3: 100 000 voxels x 300 scans, bd(nb_iter=100), FP32."""
    import pybold_b200 as pb
    V, T = 100000, 300
    y = torch.as_tensor(gen_voxels_chunked(V, T), device="cuda")
    x, z, dz, h, d = pb.bd(y, 1.0, lbda=1.7, theta_0=2.0, nb_iter=100)
    assert torch.isfinite(z).all() and torch.isfinite(d["J"]).all()
    assert int(d["n_trace"].min()) == 102 and int(d["n_trace"].max()) == 102
    assert float(d["theta"].min()) >= 0.6 and float(d["theta"].max()) <= 1.9
    # (a) same voxels, reversed order, odd-sized sub-batch: bit-identical answers
    idx = torch.arange(V - 1, V - 4098, -1, device="cuda")
    x2, z2, dz2, h2, d2 = pb.bd(y[idx].contiguous(), 1.0, lbda=1.7, theta_0=2.0, nb_iter=100)
    assert torch.equal(z2, z[idx]) and torch.equal(h2, h[idx]) and torch.equal(d2["J"], d["J"][idx])
    # (c) a sample against the CPU oracle (FP32 tolerance)
    for v in (0, 49999, V - 1):
        xo, zo, wo, ho, do = orc.bd(y[v].double().cpu().numpy(), 1.0, lbda=1.7, theta_0=2.0,
                                    nb_iter=100, theta_solver="exact")
        assert rel(z[v].cpu().numpy(), zo) < 1e-4
        assert rel(h[v].cpu().numpy(), ho) < 1e-4
        assert rel(d["J"][v].cpu().numpy(), do["J"]) < 1e-4


def test_cfg4_share_fp32_vs_fp64_and_permutation():
    """cfg4 shape (T = 1200, TR = 0.72, K = 28): one GPU's 1/8 share is 28 750 voxels."""
    import pybold_b200 as pb
    V, T, t_r = 28750, 1200, 0.72
    y = torch.as_tensor(gen_voxels_chunked(V, T, t_r), device="cuda")
    x, z, dz, h, d = pb.bd(y, t_r, lbda=1.7, theta_0=2.0, nb_iter=100)
    assert torch.isfinite(z).all()
    sub = torch.tensor([5, 17000, V - 1, 3], device="cuda")
    x64, z64, dz64, h64, d64 = pb.bd(y[sub].double(), t_r, lbda=1.7, theta_0=2.0, nb_iter=100)
    for i, v in enumerate(sub.tolist()):
        assert abs(float(d["theta"][v]) - float(d64["theta"][i])) < 1e-4
        assert rel(d["J"][v].cpu().numpy(), d64["J"][i].cpu().numpy()) < 1e-4
        assert rel(z[v].cpu().numpy(), z64[i].cpu().numpy()) < 1e-4
    x2, z2, _, _, _ = pb.bd(y[sub].contiguous(), t_r, lbda=1.7, theta_0=2.0, nb_iter=100)
    assert torch.equal(z2, z[sub])


def test_operators_stay_adjoint_at_whole_brain_size():
    """<A x, y> == <x, A^T y> on 230 000 voxels x 1200 scans (cfg4 size), FP32."""
    import pybold_b200 as pb
    V, T, K = 230000, 1200, 28
    g = torch.Generator(device="cuda").manual_seed(0)
    xx = torch.randn((V, T), device="cuda", generator=g)
    yy = torch.randn((V, T), device="cuda", generator=g)
    k = torch.rand(K, device="cuda", generator=g)
    H = pb.ConvAndLinear(pb.DiscretInteg(), k, dim_in=T)
    lhs = (H.op(xx).double() * yy.double()).sum(dim=1)
    rhs = (xx.double() * H.adj(yy).double()).sum(dim=1)
    scale = lhs.abs().max()
    assert float((lhs - rhs).abs().max() / scale) < 5e-4
