#!/usr/bin/env python
"""Benchmark of the hot path: voxels/s of the semi-blind solve ``bd()`` at 300 TRs.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--scaling weak|strong]

Headline workload (BASELINE.json configs[2], the configuration the metric is quoted on): per GPU
100 000 synthetic voxels x 300 scans, TR = 1 s, hrf_dur = 20 s (K = 20 taps),
``bd(lbda=1.7, theta_0=2.0, bounds=[(0.6, 1.9)], nb_iter=100)`` (ICASSP-2019 settings,
examples/icassp_2019/simulation.py:56-57), FP32 arithmetic.  One "step" = one pass of the whole
solve over the batch = ONE persistent kernel launch per rank, followed (N > 1) by the gather of
EVERY output (x, z, diff_z, h, theta, J, r, g) over NCCL.  Voxels are independent:
``--scaling weak`` (default) gives every rank its own 100 000 voxels, ``--scaling strong`` splits a
fixed total by voxel range.

The line printed by rank 0 follows the driver's contract; extra objects:
  roofline      FP32-FMA-pipe roofline of the solver kernel: algorithmic flops over the
                kernel's CUDA-event time, against the nominal FP32 peak at the sampled max clock
                (``frac``) and against the best FFMA microbenchmark of the same run
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, joblib over all host cores) on a bounded sample
                of the same voxels -- or the oracle port when baseline/_ref is absent
  e2e           same metric through ``pybold_b200.bd`` with pinned HOST tensors in and out
  extra         device-timed records of the other BASELINE.json configurations: cfg4 (230 000 x 1200,
                the north-star target, split over the ranks), cfg2, cfg5 and the FP64 build on cfg3
"""
from __future__ import annotations

import argparse
import gc
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxels/sec for bd() at 300 TRs"
WORKLOAD = dict(voxels_per_gpu=100000, n_scans=300, t_r=1.0, hrf_dur=20.0, lbda=1.7, theta_0=2.0,
                bounds=(0.6, 1.9), nb_iter=100)
CFG4 = dict(voxels_total=230000, n_scans=1200, t_r=0.72, hrf_dur=20.0, lbda=1.7, theta_0=2.0,
            bounds=(0.6, 1.9), nb_iter=100)
CFG2 = dict(voxels=10000, n_scans=300, t_r=1.0, hrf_dur=20.0, lbda=1.0, nb_iter=200)
CFG5 = dict(voxels=20000, n_scans=600, t_r=1.0, hrf_dur=20.0, n_lbda=64, lbda_lo=0.05, lbda_hi=20.0,
            nb_iter=200)


def hrf_taps_count(t_r, dur, dt=0.001):
    return len(range(0, int(float(dur) / dt), int(t_r / dt)))


def workload_name(kind, V, T, world=1, scaling="weak"):
    v = "%dk" % (V // 1000) if V % 1000 == 0 else str(V)
    name = "%s_%s_voxels_x_%d_TRs" % (kind, v, T)
    return name if world == 1 else "%s_per_gpu_%s" % (name, scaling) if scaling == "weak" else name + "_split"


GATHER_KEYS = {"all": ("x", "z", "diff_z", "h", "theta", "J", "r", "g"), "estimates": ("theta", "h", "z"),
               "none": ()}


def config_dict(w, world, scaling, gather="all", overlap=True, via="peer"):
    """Same keys and values in both arms (the driver compares them)."""
    V = w["voxels_per_gpu"]
    total = V * world if scaling == "weak" else V
    per_gpu = V if scaling == "weak" else -(-V // world)
    T = w["n_scans"]
    K = hrf_taps_count(w["t_r"], w["hrf_dur"])
    return {"workload": workload_name("bd", per_gpu if scaling == "weak" else total, T, world, scaling),
            "voxels_per_gpu": per_gpu, "voxels_total": total, "n_scans": T, "hrf_taps": K,
            "t_r": w["t_r"], "hrf_dur": w["hrf_dur"], "nb_iter": w["nb_iter"], "lbda": w["lbda"],
            "theta_0": w["theta_0"], "bounds": list(w["bounds"]), "scaling": scaling,
            "partition": "contiguous voxel ranges over %d rank(s)" % world,
            "gather": ("%s (%s) after every step%s" % (
                gather, ", ".join(GATHER_KEYS[gather]),
                ", on a side stream while the next step solves" if overlap and GATHER_KEYS[gather] else "")
                       if world > 1 else "single rank, nothing to gather"),
            "gather_via": ("none" if world == 1 or not GATHER_KEYS[gather] else
                           {"peer": "copy-engine writes into the peers' result tensors (CUDA IPC over NVLink), "
                                    "one-element all-reduce as the fence",
                            "nccl": "all_gather_into_tensor per output"}[via]),
            "l2": "GPU arm: L2 flushed between timed steps (256 MB write, excluded from the step time); "
                  "the working set of a step also exceeds the 126 MB L2"}


# ---- algorithmic work -----------------------------------------------------------
def mac_count(T, K):
    return T * K - K * (K - 1) // 2


def flops_bd_voxel(T, K, n, skip_tap0=False):
    """SURVEY.md 8(d), counted conservatively: the Lipschitz constant and the Newton evaluations of the
    theta step are executed but not counted (their count depends on our algorithm); ``skip_tap0``
    drops the MACs of tap 0, which is identically zero for the dilated SPM HRF and which the kernels
    do not execute."""
    mac = mac_count(T, K)
    if skip_tap0:
        mac -= T
    f_it = 4 * mac + 11 * T            # one prox-gradient iteration
    f_j = 2 * mac + 6 * T              # cost evaluation, once per outer iteration
    f_mom = 4 * mac + 2 * T            # theta step: Z^T y and the autocorrelation of z
    return (n + 1) * n * f_it + (n + 1) * f_j + n * f_mom


def flops_deconv_voxel(T, K, n):
    """``n (F_it + F_J)``; the power iteration is amortised to 0 for a shared HRF."""
    mac = mac_count(T, K)
    return n * ((4 * mac + 11 * T) + (2 * mac + 6 * T))


def bytes_bd_voxel(T, K, n, esize=4):
    return esize * (T + 3 * T + K + 1 + 3 * (n + 2)) + 4


def bytes_deconv_voxel(T, n, esize=4):
    return esize * (T + 3 * T + n) + 4


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


# ---- CPU arms ------------------------------------------------------------------------------------
def load_numpy_generator():
    """pybold_b200/synth.py's NumPy generator loaded as a plain file: the reference arm must not import
    the product package (that would load libpybold_b200.so into the reference's process)."""
    spec = importlib.util.spec_from_file_location("_pb_synth_numpy", os.path.join(ROOT, "pybold_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def oracle_bd_one(yv, t_r, lbda, theta_0, hrf_dur, bounds, nb_iter):
    from oracle import pybold_oracle as orc
    out = orc.bd(yv, t_r, lbda=lbda, theta_0=theta_0, hrf_dur=hrf_dur, bounds=[tuple(bounds)],
                 nb_iter=nb_iter, theta_solver="lbfgsb")
    return float(out[4]["J"][-1])


def cpu_bd_runner():
    """(kind, function(yv, ...) -> float): the live reference when baseline/_ref is there, else the port."""
    from baseline import reference_runner as rr
    if rr.available():
        return "reference", rr.bd_one
    return "port", oracle_bd_one


def cpu_bd_rate(pool, fn, y_sample, w):
    from joblib import delayed
    t0 = time.perf_counter()
    pool(delayed(fn)(yv, w["t_r"], w["lbda"], w["theta_0"], w["hrf_dur"], tuple(w["bounds"]), w["nb_iter"])
         for yv in y_sample)
    dt = time.perf_counter() - t0
    return len(y_sample) / dt, dt


def run_reference(args):
    """``--impl reference``: the UNMODIFIED reference's ``bd`` (baseline/_ref) on the host cores, the
    reference's own joblib fan-out over voxels, each step a bounded sample (one voxel per worker) of the
    same synthetic batch.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from baseline import reference_runner as rr
    rr.pin_threads()
    import numpy as np
    from joblib import Parallel

    synth = load_numpy_generator()
    w = dict(WORKLOAD, voxels_per_gpu=args.voxels, n_scans=args.scans, t_r=args.t_r, nb_iter=args.nb_iter)
    kind, fn = cpu_bd_runner()
    cores = os.cpu_count() or 1
    n_jobs = min(cores, 64)
    per_step = n_jobs
    warm = max(args.warmup, 1)
    y = synth.gen_voxels_chunked(per_step * (args.steps + warm), w["n_scans"], w["t_r"], w["hrf_dur"],
                                 dtype=np.float64)
    with Parallel(n_jobs=n_jobs) as pool:
        for i in range(warm):                        # Numba JIT of every worker happens here, untimed
            cpu_bd_rate(pool, fn, y[i * per_step:(i + 1) * per_step], w)
        t0 = time.perf_counter()
        for i in range(warm, warm + args.steps):
            cpu_bd_rate(pool, fn, y[i * per_step:(i + 1) * per_step], w)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d voxels per step (one per worker), %d steps, %s" % (
        per_step, args.steps,
        "unmodified pybold.bold_signal.bd from baseline/_ref" if kind == "reference"
        else "oracle port of pybold.bold_signal.bd (baseline/_ref absent)")
    line = {
        "impl": "reference", "metric": METRIC.replace("300", str(w["n_scans"])), "value": value,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(w, args.gpus, args.scaling, args.gather, not args.no_overlap, args.gather_via),
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": n_jobs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- GPU arm -------------------------------------------------------------------------------------
_RESULT_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on rank 0's stdout.  Libraries write there too (NCCL prints its version
    banner on stdout when NCCL_DEBUG is set in the environment): keep a private handle on the real stdout for
    the result line and point file descriptor 1 at stderr for everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--voxels", type=int, default=WORKLOAD["voxels_per_gpu"],
                    help="voxels per GPU (weak scaling) or in total (strong scaling)")
    ap.add_argument("--scans", type=int, default=WORKLOAD["n_scans"])
    ap.add_argument("--t-r", type=float, default=WORKLOAD["t_r"])
    ap.add_argument("--nb-iter", type=int, default=WORKLOAD["nb_iter"])
    ap.add_argument("--gather", default="all", choices=["all", "estimates", "none"],
                    help="outputs gathered over NCCL after every step when N > 1")
    ap.add_argument("--gather-via", default="peer", choices=["peer", "nccl"],
                    help="peer: copy-engine writes into the peers' result tensors (CUDA IPC over NVLink) + a one-"
                         "element all-reduce as the fence; nccl: all_gather_into_tensor per output")
    ap.add_argument("--overlap-nccl", action="store_true",
                    help="overlap the gather with the next step even on the NCCL path (measurement only)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="gather on the solver's stream instead of overlapping it with the next step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from baseline import reference_runner as rr
    import numpy as np
    import torch
    import torch.distributed as dist

    import pybold_b200 as pb
    from pybold_b200 import _lib
    from pybold_b200.bold_signal import bd_alloc, bd_batch, deconv_batch
    from pybold_b200.sharding import ALL_OUTPUTS, PeerGather, gather_outputs, voxel_range
    from pybold_b200.synth import gen_voxels_chunked, gen_voxels_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pybold_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    f32 = torch.float32

    w = dict(WORKLOAD, voxels_per_gpu=args.voxels, n_scans=args.scans, t_r=args.t_r, nb_iter=args.nb_iter)
    cfg = config_dict(w, world, args.scaling, args.gather, not args.no_overlap, args.gather_via)
    T, n, K = w["n_scans"], w["nb_iter"], cfg["hrf_taps"]
    V_total = cfg["voxels_total"]
    lo, hi = voxel_range(V_total, rank, world)
    V = hi - lo
    keys = GATHER_KEYS[args.gather]
    assert GATHER_KEYS["all"] == tuple(ALL_OUTPUTS)

    # this rank's voxel range of the global synthetic batch (seeded by global voxel index)
    y_host = torch.from_numpy(gen_voxels_chunked(V, T, w["t_r"], w["hrf_dur"], first_voxel=lo,
                                                 dtype=np.float32)).pin_memory()
    y_dev = y_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=f32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    side = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None

    def timed_steps(launch, gather, steps, warmup, nbuf=1):
        """W untimed steps, then `steps` timed ones: CUDA events on the launching stream around the whole
        loop (minus the L2 flushes between steps), per-step events around the solver launch and around the
        gather.  ``launch(b)`` solves into output buffer b, ``gather(out, b)`` gathers it.  With two buffers
        and N > 1 the gather of step i runs on a high-priority side stream while step i + 1 solves (the
        solver pulls its tasks from a work queue, so the SMs NCCL occupies for a few ms cost next to
        nothing); the last gather is exposed.  Returns (ms per step, solver ms, gather ms), MAX over ranks."""
        main = torch.cuda.current_stream()
        overlap = side is not None and nbuf > 1
        gc.collect()
        gc.disable()     # a generation-2 pass of the collector is a 10-40 ms host pause (as timeit does)
        out = None
        for _ in range(warmup):
            out = launch(0)      # held until the next call returns, like in the timed loop: a launch that
            gather(out, 0)       # allocates its outputs then finds the allocator in its steady state
            flush.fill_(1.0)
        barrier()
        marks, gmarks = [], []
        gather_done = [None] * nbuf
        t0, t1 = ev(), ev()
        t0.record()
        for i in range(steps):
            b = i % nbuf
            m = [ev() for _ in range(4)]
            if overlap and gather_done[b] is not None:
                main.wait_event(gather_done[b])      # buffer b and its gathered copies are free again
            m[0].record()
            out = launch(b)
            m[1].record()
            g0, g1 = ev(), ev()
            if overlap:
                side.wait_event(m[1])
                with torch.cuda.stream(side):
                    g0.record()
                    gather(out, b)
                    g1.record()
                gather_done[b] = g1
            else:
                g0.record()
                gather(out, b)
                g1.record()
            m[2].record()
            flush.fill_(1.0)           # L2 flush between timed iterations (excluded from the step time)
            m[3].record()
            marks.append(m)
            gmarks.append((g0, g1))
        if overlap:
            main.wait_stream(side)
        t1.record()
        barrier()
        gc.enable()
        total = t0.elapsed_time(t1)
        flush_ms = sum(m[2].elapsed_time(m[3]) for m in marks)
        kern = [m[0].elapsed_time(m[1]) for m in marks]
        gath = [a.elapsed_time(b_) for a, b_ in gmarks]
        step_ms, kern_ms, gath_ms = max_over_ranks([(total - flush_ms) / steps, sum(kern) / steps,
                                                    sum(gath) / steps])
        return step_ms, kern_ms, gath_ms, kern

    # ---- headline: device-resident inputs, reusable outputs: a step = one kernel launch (+ gather) ----
    lbda_dev = torch.full((1,), w["lbda"], dtype=f32, device=dev)
    theta0_dev = torch.full((1,), w["theta_0"], dtype=f32, device=dev)
    def out_spec(T_, K_, n_, dt):
        return {"x": ((T_,), dt), "z": ((T_,), dt), "diff_z": ((T_,), dt), "h": ((K_,), dt), "theta": ((), dt),
                "J": ((n_ + 2,), dt), "r": ((n_ + 2,), dt), "g": ((n_ + 2,), dt)}

    # The gather overlaps the next step only through the peer (copy-engine) path: an NCCL all-gather kernel on
    # a side stream competes with the persistent solve for SMs and was measured 8.7 % slower per step on
    # 8 GPUs than not overlapping at all (profiles/r02_scaling.txt); --overlap-nccl forces it for measurements.
    nbuf = 2 if (world > 1 and keys and not args.no_overlap) else 1
    peer = None
    gather_via = args.gather_via if (world > 1 and keys) else "none"
    if gather_via == "peer":
        try:        # raises on every rank or on none
            peer = [PeerGather({k: v for k, v in out_spec(T, K, n, f32).items() if k in keys}, V_total, dev)
                    for _ in range(nbuf)]
        except RuntimeError as exc:
            peer, gather_via = None, "nccl (%s)" % exc
    if peer is None and not args.overlap_nccl:
        nbuf = 1
    out_buf = [bd_alloc(V, T, K, n, f32, dev) for _ in range(nbuf)]
    gathered = [{} for _ in range(nbuf)]

    def launch(b):
        return bd_batch(y_dev, w["t_r"], lbda_dev, theta0_dev, None, w["hrf_dur"], [w["bounds"]],
                        n, False, 4, 1.0e-12, out=out_buf[b])

    def gather(out, b):
        if world > 1 and keys:   # final gather of the outputs; never inside the solve
            if peer is not None:
                peer[b].gather(out, keys)
            else:
                gather_outputs(out, V_total, keys, into=gathered[b])

    warm = max(args.warmup, 3)
    for i in range(warm):
        gather(launch(i % nbuf), i % nbuf)
    barrier()

    # ---- FP32 FMA microbenchmark (second roofline denominator), same run, same clocks regime ----
    sink = torch.empty(sms * 4 * 256, dtype=f32, device=dev)
    fma_iters = 1 << 15
    _lib.check(_lib.lib.pb_bench_fma_f32(sink.data_ptr(), sms * 4, 1 << 10, 0), "pb_bench_fma_f32")
    torch.cuda.synchronize()
    fma_tflops = 0.0
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record()
        _lib.check(_lib.lib.pb_bench_fma_f32(sink.data_ptr(), sms * 4, fma_iters,
                                             torch.cuda.current_stream().cuda_stream), "pb_bench_fma_f32")
        e1.record()
        torch.cuda.synchronize()
        fma_tflops = max(fma_tflops, sms * 4 * 256 * 16 * fma_iters * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, kern_ms, gath_ms, solver_ms = timed_steps(launch, gather, args.steps, 0, nbuf)
    clocks = sampler.stop() if rank == 0 else None
    value = V_total / (step_ms * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    nominal = sms * 128 * 2 * sm_max * 1e6 / 1e12

    def fp32_roofline(kind, v_rank, T_, K_, n_, ms, variant):
        fl = (flops_bd_voxel(T_, K_, n_) if kind == "bd" else flops_deconv_voxel(T_, K_, n_))
        by = (bytes_bd_voxel(T_, K_, n_) if kind == "bd" else bytes_deconv_voxel(T_, n_)) * v_rank
        ach = fl * v_rank / (ms * 1e-3) / 1e12
        r = {"bound": "fp32", "achieved": ach, "peak": nominal, "unit": "TFLOP/s", "frac": ach / nominal,
             "peak_source": "nominal FP32 pipe: %d SMs x 128 lanes x 2 x %.0f MHz (max SM clock sampled in this "
                            "run); MEASURED_PEAKS.json holds no FP32 figure" % (sms, sm_max),
             "peak_microbench": fma_tflops, "frac_of_microbench": ach / fma_tflops,
             "microbench": "pb_bench_fma_f32: 16 independent FFMA chains, 32 warps/SM, best of 3, this run",
             "kernel": variant, "kernel_ms": ms, "flops_per_voxel": fl,
             "hbm": {"achieved_gbs": by / (ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                     "peak_source": "measured" if "hbm_gbs" in peaks else "fallback", "bytes_per_launch": by}}
        if kind == "bd":
            fl0 = flops_bd_voxel(T_, K_, n_, skip_tap0=True)
            r["flops_per_voxel_without_tap0"] = fl0
            r["frac_without_tap0"] = fl0 * v_rank / (ms * 1e-3) / 1e12 / nominal
        return r

    def ncu_traffic(tag):
        """dram read+write bytes per launch from the committed `ncu --set full` capture of this workload
        (profiles/ncu_traffic.json: {"<tag>": {"bytes": ..., "source": "<file>"}}), else null."""
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(tag)
        except (OSError, ValueError):
            rec = None
        return (rec["bytes"], rec["source"]) if rec else (None, None)

    # ---- e2e: public API, pinned host tensors in, host tensors out ----
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            return pb.bd(y_host, w["t_r"], lbda=w["lbda"], theta_0=w["theta_0"], hrf_dur=w["hrf_dur"],
                         bounds=[w["bounds"]], nb_iter=n)
        gc.collect()
        gc.disable()
        for _ in range(3):   # untimed calls: the pinned result blocks of consecutive calls (the caller still
            res = step_e2e()  # holds the previous result) are then cached; the third call was still 60 ms slow
        barrier()
        e2e_samples = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t_s = time.perf_counter()
            res = step_e2e()                 # returns host arrays: the call ends when they are complete
            e2e_samples.append((time.perf_counter() - t_s) * 1e3)
        barrier()
        (e2e_s,) = max_over_ranks([(time.perf_counter() - t0) / args.steps])
        gc.enable()
        d2h = sum(int(a.numel()) * a.element_size() for a in res[:4]) + \
            sum(int(res[4][k].numel()) * res[4][k].element_size() for k in ("J", "r", "g", "theta", "n_trace"))
        e2e = {"value": V_total / e2e_s, "unit": "voxels/s",
               "h2d_bytes_per_step": int(y_host.numel()) * 4, "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_s * 1e3, "ms_per_step_rank0": [round(v, 2) for v in e2e_samples]}
        del res

    # ---- extra: the other BASELINE.json configurations, device-timed --------------------------------
    extra = {}
    launches = args.steps
    if not args.no_extra:
        del out_buf, y_dev, gathered      # (the peer-gather tensors stay: other ranks hold IPC mappings of them)
        torch.cuda.empty_cache()
        x_steps, x_warm = 2, 1

        # cfg4, the north-star target: 230 000 x 1200 split by voxel range over the ranks, every output gathered
        c4 = CFG4
        K4 = hrf_taps_count(c4["t_r"], c4["hrf_dur"])
        lo4, hi4 = voxel_range(c4["voxels_total"], rank, world)
        V4, T4, n4 = hi4 - lo4, c4["n_scans"], c4["nb_iter"]
        y4 = gen_voxels_device(V4, T4, c4["t_r"], c4["hrf_dur"], seed=4, first_voxel=lo4, dtype=f32)
        nb4 = 2 if (world > 1 and not args.no_overlap and (peer is not None or args.overlap_nccl)) else 1
        o4 = [bd_alloc(V4, T4, K4, n4, f32, dev) for _ in range(nb4)]
        g4 = [{} for _ in range(nb4)]

        def launch4(b):
            return bd_batch(y4, c4["t_r"], lbda_dev, theta0_dev, None, c4["hrf_dur"], [c4["bounds"]],
                            n4, False, 4, 1.0e-12, out=o4[b])

        peer4 = None
        if world > 1 and peer is not None:
            peer4 = [PeerGather(out_spec(T4, K4, n4, f32), c4["voxels_total"], dev) for _ in range(nb4)]

        def gather4(out, b):
            if world > 1:
                if peer4 is not None:
                    peer4[b].gather(out)
                else:
                    gather_outputs(out, c4["voxels_total"], ALL_OUTPUTS, into=g4[b])

        s_ms, k_ms, g_ms, _ = timed_steps(launch4, gather4, x_steps, x_warm, nb4)
        launches += x_steps
        (v4max,) = max_over_ranks([float(V4)])
        extra["cfg4_bd_230k_x_1200"] = {
            "workload": workload_name("bd", c4["voxels_total"], T4, world, "strong"), "dtype": "f32",
            "value": c4["voxels_total"] / (s_ms * 1e-3), "unit": "voxels/s", "ms_per_step": s_ms,
            "solver_ms": k_ms, "gather_ms": g_ms, "gather_exposed_ms": max(s_ms - k_ms, 0.0),
            "gather_overlapped": nb4 > 1, "steps": x_steps, "warmup": x_warm,
            "voxels_total": c4["voxels_total"], "voxels_this_rank_max": int(v4max), "n_scans": T4,
            "hrf_taps": K4, "nb_iter": n4, "data": "synthetic (device generator philox-v1, seed 4)",
            "roofline": fp32_roofline("bd", int(v4max), T4, K4, n4, k_ms,
                                      "variant %d" % _lib.lib.pb_solver_variant(T4, K4, 0))}
        del y4, o4, g4, peer4
        torch.cuda.empty_cache()

        if world == 1:
            # cfg3 on the FP64 build (what float64 input, the reference's type, runs)
            f64 = torch.float64
            y64 = y_host.to(dev).to(f64)
            o64 = bd_alloc(V, T, K, n, f64, dev)
            lb64, th64 = lbda_dev.to(f64), theta0_dev.to(f64)
            s_ms, k_ms, _, _ = timed_steps(
                lambda b: bd_batch(y64, w["t_r"], lb64, th64, None, w["hrf_dur"], [w["bounds"]], n, False, 4,
                                   1.0e-12, out=o64), lambda out, b: None, x_steps, x_warm)
            launches += x_steps
            fl = flops_bd_voxel(T, K, n) * V / (k_ms * 1e-3) / 1e12
            extra["cfg3_bd_fp64_build"] = {
                "workload": workload_name("bd", V, T), "dtype": "f64", "value": V / (s_ms * 1e-3),
                "unit": "voxels/s", "ms_per_step": s_ms, "steps": x_steps, "warmup": x_warm,
                "roofline": {"bound": "fp64", "achieved": fl, "unit": "TFLOP/s",
                             "note": "parity build; the FP64 pipe of B200 is not a roofline this path is held to",
                             "kernel": "variant %d" % _lib.lib.pb_solver_variant(T, K, 1), "kernel_ms": k_ms}}
            del y64, o64
            torch.cuda.empty_cache()

            # cfg2: known-HRF deconv batch, 10 000 x 300, fixed lambda, 200 iterations, shared HRF
            c2 = CFG2
            K2 = hrf_taps_count(c2["t_r"], c2["hrf_dur"])
            y2 = gen_voxels_device(c2["voxels"], c2["n_scans"], c2["t_r"], c2["hrf_dur"], seed=2, dtype=f32)
            h2 = torch.as_tensor(pb.spm_hrf(1.0, c2["t_r"], c2["hrf_dur"], True)[0], dtype=f32, device=dev)
            x0 = np.random.RandomState(0).randn(c2["n_scans"])
            L2c = 0.9 * pb.utils.spectral_radius_est(pb.ConvAndLinear(pb.DiscretInteg(), h2, dim_in=c2["n_scans"]),
                                                     (c2["n_scans"],), x0=x0.astype(np.float32))
            L2t = torch.full((1,), float(L2c), dtype=f32, device=dev)
            lb2 = torch.full((1,), c2["lbda"], dtype=f32, device=dev)
            k_evs = []

            def launch2(b):
                k_evs.append((ev(), ev()))
                return deconv_batch(y2, h2, lb2, L2t, None, False, 1.0e-6, 6, c2["nb_iter"], events=k_evs[-1])

            s_ms, _, _, each2 = timed_steps(launch2, lambda out, b: None, 5, 3)
            # the launch is so short that events around the Python call would time the host: these two sit
            # immediately around the C-ABI launch (the momentum-table kernel is part of the launch)
            k_ms = sum(a.elapsed_time(b) for a, b in k_evs[-5:]) / 5
            launches += 5
            extra["cfg2_deconv_10k_x_300"] = {
                "workload": workload_name("deconv", c2["voxels"], c2["n_scans"]), "dtype": "f32",
                "value": c2["voxels"] / (s_ms * 1e-3), "unit": "voxels/s", "ms_per_step": s_ms, "steps": 5,
                "warmup": 3, "nb_iter": c2["nb_iter"], "lbda": c2["lbda"],
                "ms_each_call": [round(v, 3) for v in each2],   # events around each Python call (rank 0)
                "note": "ms_per_step is the Python-level call (output allocation, host launch latency); the roofline "
                        "uses events placed around the kernel launch; 10 000 voxels are %.1f waves of the "
                        "grid: tail-bound" % (c2["voxels"] / max(1, sms * 24)),
                "roofline": fp32_roofline("deconv", c2["voxels"], c2["n_scans"], K2, c2["nb_iter"], k_ms,
                                          "fast_deconvg_kernel")}
            del y2
            torch.cuda.empty_cache()

            # cfg5: regularisation path, 64 lambdas x 20 000 voxels x 600 scans, 200 iterations
            c5 = CFG5
            K5 = hrf_taps_count(c5["t_r"], c5["hrf_dur"])
            y5 = gen_voxels_device(c5["voxels"], c5["n_scans"], c5["t_r"], c5["hrf_dur"], seed=5, dtype=f32)
            h5 = torch.as_tensor(pb.spm_hrf(1.0, c5["t_r"], c5["hrf_dur"], True)[0], dtype=f32, device=dev)
            lbdas = np.geomspace(c5["lbda_lo"], c5["lbda_hi"], c5["n_lbda"])
            x05 = np.random.RandomState(0).randn(c5["n_scans"]).astype(np.float32)
            from pybold_b200.bold_signal import deconv_lbda_path
            problems = c5["n_lbda"] * c5["voxels"]

            def path(b):
                return deconv_lbda_path(y5, c5["t_r"], h5, lbdas, nb_iter=c5["nb_iter"], x0=x05)

            s_ms, k_ms, _, each5 = timed_steps(path, lambda out, b: None, 2, 2)
            launches += 2
            extra["cfg5_lbda_path_64_x_20k_x_600"] = {
                "workload": "deconv_path_%d_lbda_x_%dk_voxels_x_%d_TRs" % (c5["n_lbda"], c5["voxels"] // 1000,
                                                                         c5["n_scans"]),
                "dtype": "f32", "value": problems / (s_ms * 1e-3), "unit": "(lambda, voxel) problems/s",
                "ms_per_step": s_ms, "steps": 2, "warmup": 2, "nb_iter": c5["nb_iter"],
                "ms_each_call": [round(v, 3) for v in each5],
                "note": "timed through deconv_lbda_path (public API): power iteration, launches, output "
                        "allocation and the J normalisation are inside ms_per_step",
                "roofline": fp32_roofline("deconv", problems, c5["n_scans"], K5, c5["nb_iter"], k_ms,
                                          "fast_deconv_kernel (whole path call)")}
            del y5
            torch.cuda.empty_cache()

    if rank == 0:
        vid = _lib.lib.pb_solver_variant(T, K, 0)
        roofline = fp32_roofline("bd", V, T, K, n, kern_ms, "%s (variant %d = lanes per voxel x 1e6 + samples per "
                                 "lane x 1e3 + unrolled taps)" % ("fast_bdg_kernel" if vid // 1000000 <= 16 else
                                                                   "fast_bdc_kernel" if vid else "generic_bd_kernel", vid))
        roofline["kernel_ms_per_step"] = solver_ms
        tr_bytes, tr_src = ncu_traffic(cfg["workload"])
        roofline["traffic"] = tr_bytes
        roofline["traffic_source"] = tr_src
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            from joblib import Parallel
            rr.pin_threads()        # one BLAS / OpenMP thread per worker; the workers are spawned after this
            kind, fn = cpu_bd_runner()
            cores = min(os.cpu_count() or 1, 64)
            y_s = y_host[:cores * 5].numpy().astype(np.float64)
            with Parallel(n_jobs=cores) as pool:
                cpu_bd_rate(pool, fn, y_s[:cores], w)           # warm-up: Numba JIT in every worker
                rate, dt = cpu_bd_rate(pool, fn, y_s[cores:], w)
            cpu = {"value": rate, "unit": "voxels/s", "cores": cores, "kind": kind,
                   "sample": "voxels %d..%d of the same batch (after one warm-up voxel per worker), %.1f s wall, "
                             "joblib n_jobs=%d, %s" % (cores, len(y_s) - 1, dt, cores,
                                                       "unmodified pybold.bold_signal.bd (baseline/_ref)"
                                                       if kind == "reference" else "oracle port")}
        line = {
            "metric": METRIC.replace("300", str(T)), "value": value, "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "timing": {"solver_ms": kern_ms, "gather_ms": gath_ms,
                       "gather_exposed_ms": max(step_ms - kern_ms, 0.0), "gather_overlapped": nbuf > 1,
                       "gather_path": gather_via,
                       "note": "CUDA events, max over ranks; solver_ms around the solver launch on its stream; "
                               "gather_ms around the gather calls on the stream they run on -- overlapped, its fence "
                               "only completes when the next solve lets a CTA in, what a step pays is "
                               "gather_exposed_ms = ms_per_step - solver_ms"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu, "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
