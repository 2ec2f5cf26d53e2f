"""Secondary measurements (not the driver's contract): the other BASELINE.json configs.

    python tools/bench_configs.py [cfg2] [cfg3_native] [cfg4] [cfg5] [f64]

Prints one JSON line per configuration: voxels/s (device-resident inputs, CUDA events, best of
3 launches after a warm-up) and the algorithmic Tflop/s of SURVEY.md 8(d).
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pybold_b200 import _lib  # noqa: E402
from pybold_b200.bold_signal import bd_alloc, bd_batch, deconv_batch  # noqa: E402
from pybold_b200.hrf_model import hrf_len, spm_hrf  # noqa: E402
from pybold_b200.synth import gen_voxels_chunked  # noqa: E402
from pybold_b200.utils import spectral_radius_est  # noqa: E402
from pybold_b200.linear import ConvAndLinear, DiscretInteg  # noqa: E402


def mac(T, K):
    return T * K - K * (K - 1) // 2


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run_bd(name, V, T, t_r, n=100, dtype=torch.float32):
    K = hrf_len(t_r, 20.0)
    y = torch.as_tensor(gen_voxels_chunked(V, T, t_r, 20.0), device="cuda", dtype=dtype)
    out = bd_alloc(V, T, K, n, dtype, y.device)
    lb = torch.full((1,), 1.7, dtype=dtype, device="cuda")
    th = torch.full((1,), 2.0, dtype=dtype, device="cuda")
    ms = timed(lambda: bd_batch(y, t_r, lb, th, None, 20.0, [(0.6, 1.9)], n, False, 4, 1e-12, out=out))
    f_it, f_j, f_mom = 4 * mac(T, K) + 11 * T, 2 * mac(T, K) + 6 * T, 4 * mac(T, K) + 2 * T
    flops = V * ((n + 1) * n * f_it + (n + 1) * f_j + n * f_mom)
    print(json.dumps({"config": name, "solver": "bd", "V": V, "T": T, "K": K, "nb_iter": n,
                      "dtype": str(dtype).split(".")[-1], "ms": ms, "voxels_per_s": V / ms * 1e3,
                      "algorithmic_tflops": flops / ms / 1e9,
                      "variant": _lib.lib.pb_solver_variant(T, K, int(dtype == torch.float64))}), flush=True)


def run_deconv(name, V, T, t_r, n=200, n_lbda=1, dtype=torch.float32):
    K = hrf_len(t_r, 20.0)
    y = torch.as_tensor(gen_voxels_chunked(V, T, t_r, 20.0), device="cuda", dtype=dtype)
    h, _ = spm_hrf(1.0, t_r, 20.0, True)
    hd = torch.as_tensor(h, device="cuda", dtype=dtype)
    x0 = torch.as_tensor(np.random.RandomState(0).randn(T), device="cuda", dtype=dtype)
    L = 0.9 * spectral_radius_est(ConvAndLinear(DiscretInteg(), hd, T), (T,), x0=x0)
    if n_lbda > 1:      # cfg5: lambda sweep = one voxel-problem per (lambda, voxel) pair
        lbdas = torch.as_tensor(np.geomspace(0.05, 20, n_lbda), device="cuda", dtype=dtype)
        y = y.repeat(n_lbda, 1)
        lb = lbdas.repeat_interleave(V)
        Vtot = V * n_lbda
    else:
        lb, Vtot = 1.0, V
    ms = timed(lambda: deconv_batch(y, hd, lb, L, None, False, 1e-6, 6, n))
    f_it, f_j = 4 * mac(T, K) + 11 * T, 2 * mac(T, K) + 6 * T
    flops = Vtot * n * (f_it + f_j)
    print(json.dumps({"config": name, "solver": "deconv", "V": Vtot, "T": T, "K": K, "nb_iter": n,
                      "dtype": str(dtype).split(".")[-1], "ms": ms, "problems_per_s": Vtot / ms * 1e3,
                      "algorithmic_tflops": flops / ms / 1e9}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3_native", "cfg4", "cfg5", "f64"]
    if "cfg2" in which:
        run_deconv("cfg2: deconv 10k x 300, 200 it", 10000, 300, 1.0)
        run_deconv("cfg2 x10: deconv 100k x 300, 200 it", 100000, 300, 1.0)
    if "cfg3_native" in which:
        run_bd("cfg3 native ICASSP shape: bd 100k x 240, TR 0.75", 100000, 240, 0.75)
    if "cfg4" in which:
        run_bd("cfg4 (one GPU's share at 8 GPUs): bd 28750 x 1200, TR 0.72", 28750, 1200, 0.72)
    if "cfg5" in which:
        run_deconv("cfg5 (1/8 of the grid): deconv 8 lbda x 20k x 600", 20000, 600, 1.0, n_lbda=8)
    if "f64" in which:
        run_bd("cfg3 in FP64 (parity build): bd 20k x 300", 20000, 300, 1.0, dtype=torch.float64)
        run_deconv("cfg2 in FP64: deconv 10k x 300", 10000, 300, 1.0, dtype=torch.float64)
