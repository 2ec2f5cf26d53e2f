"""bench.py's contract pieces that need no GPU: both arms describe the same configuration, the algorithmic
work figures are the documented ones, the reference arm prints one valid JSON line timing the unmodified
reference (baseline/_ref) or, without it, the oracle port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_algorithmic_work_figures():
    # cfg3 (T = 300, K = 20, nb_iter = 100)
    assert bench.mac_count(300, 20) == 5810
    assert bench.flops_bd_voxel(300, 20, 100) == 271793420
    assert bench.flops_bd_voxel(300, 20, 100, skip_tap0=True) == 259492820
    assert bench.flops_deconv_voxel(300, 20, 200) == 200 * (26540 + 13420)       # 7.99 Mflop
    assert bench.bytes_bd_voxel(300, 20, 100) == 6112
    assert bench.hrf_taps_count(1.0, 20.0) == 20 and bench.hrf_taps_count(0.72, 20.0) == 28
    assert bench.hrf_taps_count(0.75, 20.0) == 27 and bench.hrf_taps_count(0.5, 20.0) == 40


def test_both_arms_print_the_same_config():
    w = dict(bench.WORKLOAD)
    for world in (1, 2, 8):
        for scaling in ("weak", "strong"):
            a = bench.config_dict(w, world, scaling, "all", True, "peer")
            b = bench.config_dict(dict(w), world, scaling, "all", True, "peer")
            assert a == b and json.dumps(a) == json.dumps(b)
            assert a["voxels_total"] == (100000 * world if scaling == "weak" else 100000)
    assert bench.config_dict(w, 1, "weak")["workload"] == "bd_100k_voxels_x_300_TRs"
    w2 = dict(w, voxels_per_gpu=28750, n_scans=1200, t_r=0.72)
    c = bench.config_dict(w2, 8, "weak")
    assert c["workload"] == "bd_28750_voxels_x_1200_TRs_per_gpu_weak" and c["hrf_taps"] == 28


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, NUMBA_CACHE_DIR="/tmp/pybold_ref_numba_cache")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--nb-iter", "6"], capture_output=True, text=True, env=env,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "voxels/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    from baseline import reference_runner as rr
    assert d["cpu_baseline"]["kind"] == ("reference" if rr.available() else "port")
    want = bench.config_dict(dict(bench.WORKLOAD, nb_iter=6), 1, "weak", "all", True, "peer")
    assert d["config"] == want
    # the reference arm must not have loaded the product library
    assert "pybold_b200" not in out.stderr


def test_stdout_carries_only_the_result_line():
    """Libraries write to file descriptor 1 (NCCL's version banner with NCCL_DEBUG set): the bench keeps a
    private handle for its JSON line and sends everything else to stderr."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version x.y\\n'); print('noise'); bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version" in r.stderr and "noise" in r.stderr
