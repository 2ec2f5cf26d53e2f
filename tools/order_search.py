"""Search the source order of the unrolled R x K tile (csrc/pb_tile.cuh) with the static model of
tools/rf_model.py -- no GPU needed.

    python tools/order_search.py group|group_deconv|cta|warp [jobs]

Compiles one instantiation of the kernel family per candidate order (tap direction, sample direction,
accumulator block size for the convolution and for the correlation) with -D overrides of the PB_*_JDESC /
_RDESC / _RB macros and prints the candidates by modelled cycles.  Time the best few on a B200
(tools/exp_bdg.cu) before changing the defaults in pb_fastg.cuh / pb_fastc.cuh / pb_fast.cuh.
"""
import itertools
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rf_model import score  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = {
    "group": ("pb_fastg.cuh", "fast_bdg_launch<float, 19, 20, 16, 8, 4, 3>", "BdArgs", "fast_bdg_kernel", "PB_"),
    "group_smh2": ("pb_fastg.cuh", "fast_bdg_launch<float, 19, 20, 16, 8, 4, 3, 0, 2>", "BdArgs", "fast_bdg_kernel", "PB_"),
    "group_smh2_m4": ("pb_fastg.cuh", "fast_bdg_launch<float, 19, 20, 16, 8, 4, 4, 0, 2>", "BdArgs", "fast_bdg_kernel", "PB_"),
    "group_deconv": ("pb_fastg.cuh", "fast_deconvg_launch<float, 19, 20, 16, 8, 4, 3>", "DeconvArgs", "fast_deconvg_kernel", "PB_"),
    "cta": ("pb_fastc.cuh", "fast_bdc_launch<float, 20, 28, 2, 6>", "BdArgs", "fast_bdc_kernel", "PB_C_"),
    "warp": ("pb_fast.cuh", "fast_deconv_launch<float, 20, 20, true, 4, 3>", "DeconvArgs", "fast_deconv_kernel", "PB_W_"),
}


def main():
    fam = sys.argv[1] if len(sys.argv) > 1 else "group"
    jobs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
    header, launch, args, kern, pre = FAMILY[fam]
    tmp = tempfile.mkdtemp(prefix="pb_order_")
    src = os.path.join(tmp, "inst.cu")
    with open(src, "w") as f:
        f.write('#include "%s"\nnamespace pb { int inst(const %s<float> &a, cudaStream_t s) { return %s(a, s); } }\n'
                % (header, args, launch))
    cands = []
    conv = [{"CONV_JDESC": cj, "CONV_RB": crb} for cj, crb in itertools.product((0, 1), (0, 5, 7, 8, 10))]
    corr = [{"CORR_JDESC": kj, "CORR_RDESC": kr, "CORR_RB": krb}
            for kj, kr, krb in itertools.product((0, 1), (0, 1), (0, 5, 7, 8))]
    if os.environ.get("PB_SEARCH_DS"):      # data-stationary orders (all taps of one window position)
        conv_ds = [{"CONV_DS": ds, "CONV_JDESC": cj} for ds, cj in itertools.product((1, 2), (0, 1))]
        corr_ds = [{"CORR_DS": ds, "CORR_JDESC": kj} for ds, kj in itertools.product((1, 2), (0, 1))]
        pairs = [(a, b) for a in conv_ds for b in corr + corr_ds] + [(a, b) for a in conv for b in corr_ds]
    else:
        pairs = [(a, b) for a in conv for b in corr]
    for a, b in pairs:
        c = dict(a)
        c.update(b)
        cands.append(c)

    def run(i):
        c = cands[i]
        out = os.path.join(tmp, "c%d.cubin" % i)
        flags = ["-D%s%s=%d" % (pre, k, v) for k, v in c.items()] + os.environ.get("PB_EXTRA_FLAGS", "").split()
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                        "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "pybold_b200", "csrc")] + flags +
                       ["-cubin", "-o", out, src], check=True, capture_output=True)
        n, reads, cyc, f3 = score(out, kern)
        return cyc, n, reads, f3, c

    with ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(run, range(len(cands))))
    for cyc, n, reads, f3, c in sorted(res, key=lambda r: r[0])[:12]:
        print("%5d modelled cycles (%d instructions, reads %s, %d FFMA with three register reads)  %s"
              % (cyc, n, reads, f3, " ".join("%s=%d" % kv for kv in c.items())))


if __name__ == "__main__":
    main()
