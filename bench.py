#!/usr/bin/env python
"""Benchmark of the hot path: voxels/s of the semi-blind solve ``bd()`` at 300 TRs.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): per GPU
100 000 synthetic voxels x 300 scans, TR = 1 s, hrf_dur = 20 s (K = 20 taps),
``bd(lbda=1.7, theta_0=2.0, bounds=[(0.6, 1.9)], nb_iter=100)`` (ICASSP-2019 settings,
examples/icassp_2019/simulation.py:56-57), FP32 arithmetic.  One "step" = one pass of the whole
solve over the batch = ONE persistent kernel launch per rank.  Voxels are independent, so
N ranks each solve their own 100 000 voxels (weak scaling) and the only collective is the final
all-gather of the estimates (theta, h, z) over NCCL.

The line printed by rank 0 follows the driver's contract; extra objects:
  roofline      FP32-FMA-pipe roofline of the solver kernel (algorithmic flops / CUDA-event time
                against an FMA microbenchmark measured in the same run) + the HBM figure
  cpu_baseline  the CPU oracle (restatement of the reference, SciPy L-BFGS-B theta step like the
                reference) on a bounded sample of the same voxels, joblib over all host cores
  e2e           same metric through ``pybold_b200.bd`` with pinned HOST tensors in and out
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="bd_100k_voxels_x_300_TRs", voxels_per_gpu=100000, n_scans=300, t_r=1.0,
                hrf_dur=20.0, lbda=1.7, theta_0=2.0, bounds=(0.6, 1.9), nb_iter=100)


def algorithmic_flops_per_voxel(T, K, n):
    """SURVEY.md 8(d), counted conservatively (see DESIGN.md "Roofline accounting")."""
    mac = T * K - K * (K - 1) // 2
    f_it = 4 * mac + 11 * T            # one prox-gradient iteration
    f_j = 2 * mac + 6 * T              # cost evaluation, once per outer iteration
    f_mom = 4 * mac + 2 * T            # theta step: Z^T y and the autocorrelation of z
    return (n + 1) * n * f_it + (n + 1) * f_j + n * f_mom


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the solver kernel, from the committed
# `ncu --set full` capture of this very command (profiles/r01_ncu_bd_t300_reordered.txt); only valid for
# the default workload, null otherwise.
NCU_TRAFFIC_BYTES = {(100000, 300, 100): 628.3e6}


def algorithmic_bytes_per_voxel(T, K, n):
    return 4 * (T + 3 * T + K + 1 + 3 * (n + 2) + 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def cpu_oracle_rate(y_sample, w, n_jobs):
    """voxels/s of the CPU restatement of the reference on ``y_sample`` (joblib over voxels)."""
    import contextlib
    import io

    from joblib import Parallel, delayed

    from oracle import pybold_oracle as orc

    def one(yv):
        with contextlib.redirect_stdout(io.StringIO()):
            orc.bd(yv, w["t_r"], lbda=w["lbda"], theta_0=w["theta_0"], hrf_dur=w["hrf_dur"],
                   bounds=[w["bounds"]], nb_iter=w["nb_iter"], theta_solver="lbfgsb")
        return 0

    t0 = time.perf_counter()
    Parallel(n_jobs=n_jobs)(delayed(one)(yv) for yv in y_sample)
    dt = time.perf_counter() - t0
    return len(y_sample) / dt, dt


def run_reference(args):
    """``--impl reference``: the reference's CPU algorithm (oracle port: the Python reference is
    not installable on the GPU box) on the host cores, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from pybold_b200.synth import gen_voxels_chunked
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"                         # like the reference's examples (validation.py:6-9)
    w = dict(WORKLOAD)
    cores = os.cpu_count() or 1
    n_jobs = min(cores, 64)
    per_step = n_jobs
    y = gen_voxels_chunked(per_step * (args.steps + args.warmup), w["n_scans"], w["t_r"], w["hrf_dur"],
                           dtype=np.float64)
    for i in range(args.warmup):
        cpu_oracle_rate(y[i * per_step:(i + 1) * per_step], w, n_jobs)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        cpu_oracle_rate(y[i * per_step:(i + 1) * per_step], w, n_jobs)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d voxels per step (one per worker), %d steps" % (per_step, args.steps)
    line = {
        "impl": "reference", "metric": "voxels/sec for bd() at 300 TRs", "value": value,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "n_scans": w["n_scans"], "nb_iter": w["nb_iter"],
                   "lbda": w["lbda"], "note": "CPU oracle port of pybold.bold_signal.bd "
                   "(dense Toeplitz/Gram matrices, SciPy L-BFGS-B theta step), joblib over voxels"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": n_jobs, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voxels", type=int, default=WORKLOAD["voxels_per_gpu"], help="voxels per GPU")
    ap.add_argument("--scans", type=int, default=WORKLOAD["n_scans"])
    ap.add_argument("--t-r", type=float, default=WORKLOAD["t_r"])
    ap.add_argument("--nb-iter", type=int, default=WORKLOAD["nb_iter"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import pybold_b200 as pb
    from pybold_b200 import _lib
    from pybold_b200.bold_signal import bd_alloc, bd_batch
    from pybold_b200.sharding import gather_rows
    from pybold_b200.synth import gen_voxels_chunked

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pybold_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = dict(WORKLOAD, voxels_per_gpu=args.voxels, n_scans=args.scans, t_r=args.t_r,
             nb_iter=args.nb_iter)
    V, T, n = w["voxels_per_gpu"], w["n_scans"], w["nb_iter"]
    K = pb.hrf_model.hrf_len(w["t_r"], w["hrf_dur"])
    V_total = V * world

    # this rank's voxel range of the global synthetic batch (seeded by global voxel index)
    y_host = torch.from_numpy(gen_voxels_chunked(V, T, w["t_r"], w["hrf_dur"], first_voxel=rank * V,
                                                 dtype=np.float32)).pin_memory()
    y_dev = y_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    # device-resident parameters and reusable output buffers: a step is then exactly one
    # asynchronous kernel launch (no allocation, no host copy inside the timed region)
    lbda_dev = torch.full((1,), w["lbda"], dtype=torch.float32, device=dev)
    theta0_dev = torch.full((1,), w["theta_0"], dtype=torch.float32, device=dev)
    out_buf = bd_alloc(V, T, K, n, torch.float32, dev)

    def launch():
        return bd_batch(y_dev, w["t_r"], lbda_dev, theta0_dev, None, w["hrf_dur"], [w["bounds"]],
                        n, False, 4, 1.0e-12, out=out_buf)

    def step_device():
        out = launch()
        if world > 1:   # final gather of the estimates; never inside the solve
            for key in ("theta", "h", "z"):
                out[key + "_all"] = gather_rows(out[key], V_total)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
        flush.fill_(1.0)
    barrier()

    # ---- FP32 FMA microbenchmark (roofline denominator), same run, same clocks regime ----
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sink = torch.empty(sms * 8 * 256, dtype=torch.float32, device=dev)
    fma_iters = 1 << 16
    _lib.check(_lib.lib.pb_bench_fma_f32(sink.data_ptr(), sms * 8, 1 << 12, 0), "pb_bench_fma_f32")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.lib.pb_bench_fma_f32(sink.data_ptr(), sms * 8, fma_iters,
                                         torch.cuda.current_stream().cuda_stream), "pb_bench_fma_f32")
    e1.record()
    torch.cuda.synchronize()
    fma_tflops = sms * 8 * 256 * 8 * fma_iters * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12

    # ---- timed region: exactly K steps, device timing, max over ranks ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms = []
    t_evt0, t_evt1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush_ms = 0.0
    t_evt0.record()
    for _ in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        out = launch()
        k1.record()
        if world > 1:
            for key in ("theta", "h", "z"):
                out[key + "_all"] = gather_rows(out[key], V_total)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        flush.fill_(1.0)           # L2 flush between timed iterations (excluded from the step time)
        f1.record()
        kernel_ms.append((k0, k1, f0, f1))
    t_evt1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_evt0.elapsed_time(t_evt1)
    flush_ms = sum(f0.elapsed_time(f1) for (_, _, f0, f1) in kernel_ms)
    solver_ms = [k0.elapsed_time(k1) for (k0, k1, _, _) in kernel_ms]
    step_ms = (total_ms - flush_ms) / args.steps
    t = torch.tensor([step_ms, sum(solver_ms) / len(solver_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(t[0]), float(t[1])
    value = V_total / (step_ms * 1e-3)

    # ---- e2e: public API, pinned host tensors in, host tensors out ----
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            return pb.bd(y_host, w["t_r"], lbda=w["lbda"], theta_0=w["theta_0"], hrf_dur=w["hrf_dur"],
                         bounds=[w["bounds"]], nb_iter=n)
        res = step_e2e()     # two untimed calls: the pinned result blocks of two consecutive calls
        res = step_e2e()     # (the caller still holds the previous result) are then both cached
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        t2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        d2h = sum(int(a.numel()) * a.element_size() for a in res[:4]) + \
            sum(int(res[4][k].numel()) * res[4][k].element_size() for k in ("J", "r", "g", "theta", "n_trace"))
        e2e = {"value": V_total / float(t2[0]), "unit": "voxels/s",
               "h2d_bytes_per_step": int(y_host.numel()) * 4, "d2h_bytes_per_step": int(d2h)}

    if rank == 0:
        flops = algorithmic_flops_per_voxel(T, K, n) * V
        hbm_bytes = algorithmic_bytes_per_voxel(T, K, n) * V
        achieved = flops / (kern_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        nominal = sms * 128 * 2 * sm_max * 1e6 / 1e12
        roofline = {
            "bound": "fp32", "achieved": achieved, "peak": fma_tflops, "unit": "TFLOP/s",
            "frac": achieved / fma_tflops, "traffic": NCU_TRAFFIC_BYTES.get((V, T, n)),
            "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_ncu_bd_t300_reordered.txt); "
                            "algorithmic %.1f MB" % (hbm_bytes / 1e6),
            "peak_source": "FFMA microbenchmark in this run (pb_bench_fma_f32); nominal "
                           "%d SMs x 128 lanes x 2 x %.0f MHz = %.1f" % (sms, sm_max, nominal),
            "frac_of_nominal": achieved / nominal,
            "kernel": "fast_bd_kernel (variant %d)" % _lib.lib.pb_solver_variant(T, K, 0),
            "kernel_ms": kern_ms, "kernel_ms_per_step": solver_ms, "flops_per_voxel": algorithmic_flops_per_voxel(T, K, n),
            "hbm": {"achieved_gbs": hbm_bytes / (kern_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                    "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                    "bytes_per_voxel": algorithmic_bytes_per_voxel(T, K, n)},
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = min(os.cpu_count() or 1, 64)
            n_s = cores * 6
            y_s = y_host[:n_s].numpy().astype(np.float64)
            rate, dt = cpu_oracle_rate(y_s, w, cores)
            cpu = {"value": rate, "unit": "voxels/s", "cores": cores, "kind": "port",
                   "sample": "first %d voxels of the same batch, %.1f s wall, joblib n_jobs=%d"
                             % (n_s, dt, cores)}
        line = {
            "metric": "voxels/sec for bd() at 300 TRs", "value": value, "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "voxels_per_gpu": V, "voxels_total": V_total,
                       "n_scans": T, "hrf_taps": K, "t_r": w["t_r"], "nb_iter": n, "lbda": w["lbda"],
                       "theta_0": w["theta_0"], "bounds": list(w["bounds"]),
                       "partition": "voxel ranges, %d rank(s), final all-gather of theta/h/z only" % world,
                       "l2": "flushed between timed steps (256 MB write, excluded from the step time); "
                             "working set per step %.0f MB > 126 MB L2" % (hbm_bytes / 1e6)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
