"""Static register-file / issue model of a SASS loop (profiles/r01_rf_bandwidth.txt).

    python tools/rf_model.py file.cubin [kernel-name-substring]

Finds the FFMA-densest loop (>= 400 instructions) of the kernel, counts per register bank (even / odd
register index) the 32-bit operand vectors its instructions read -- an operand flagged `.reuse` by the
previous instruction in the same slot is free -- and prints max(instructions, reads(even), reads(odd)):
the number of issue cycles one warp needs per loop iteration on a B200 sub-partition.  Measured kernel
times follow it at 0.236-0.241 ms per modelled cycle (tools/exp_bdg.cu, 42 624 voxels).
"""
import collections
import re
import subprocess
import sys

NODST = ("ST", "STS", "STG", "BRA", "BAR", "WARPSYNC", "EXIT", "RED", "ATOM", "BSYNC", "BSSY", "NOP", "ISETP",
         "FSETP", "DSETP", "PLOP3", "R2UR")


def kernel_sass(cubin, kern):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout.splitlines()
    on, lines = False, []
    for ln in out:
        if "Function :" in ln:
            on = kern in ln
        elif on and re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", ln):
            lines.append(ln)
    return lines


def hot_loop(lines, min_len=400):
    addr = [int(re.search(r"/\*([0-9a-f]+)\*/", ln).group(1), 16) for ln in lines]
    best = None
    for i, ln in enumerate(lines):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", ln)
        if not m:
            continue
        t = int(m.group(1), 16)
        if t >= addr[i] or (addr[i] - t) // 16 < min_len:
            continue
        n = (addr[i] - t) // 16
        nf = sum(1 for l2, a in zip(lines, addr) if t <= a <= addr[i] and "FFMA" in l2)
        if best is None or nf / n > best[0]:
            best = (nf / n, t, addr[i])
    if best is None:
        return lines
    return [ln for ln, a in zip(lines, addr) if best[1] <= a <= best[2]]


def model(lines):
    reads, n, prev = [0, 0], 0, {}
    per = collections.Counter()
    for ln in lines:
        m = re.search(r"\*/\s+(@!?U?P\d\s+)?(\S+)\s*(.*?);", ln)
        if not m:
            continue
        base = m.group(2).split(".")[0]
        srcs = [a.strip() for a in m.group(3).split(",")] if m.group(3) else []
        n += 1
        if base not in NODST and not base.startswith("ST") and srcs:
            while srcs and re.match(r"^!?U?P(T|\d)$", srcs[0]):
                srcs.pop(0)
            if srcs:
                srcs.pop(0)                      # destination register
        wide = 2 if base in ("DFMA", "DADD", "DMUL", "DSETP") else 1
        new, r_this = {}, 0
        for i, s in enumerate(srcs):
            for mm in re.finditer(r"(?<![U\w])R(\d+)(\.reuse)?", s):
                reg = int(mm.group(1))
                if prev.get(i) != reg:
                    for w in range(wide):
                        reads[(reg + w) % 2] += 1
                        r_this += 1
                if mm.group(2):
                    new[i] = reg
        prev = new
        per[(base, r_this)] += 1
    return n, reads, per


def score(cubin, kern):
    n, reads, per = model(hot_loop(kernel_sass(cubin, kern)))
    return n, reads, max(n, *reads), per.get(("FFMA", 3), 0)


if __name__ == "__main__":
    cubin = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else "fast_bdg_kernel"
    n, reads, per = model(hot_loop(kernel_sass(cubin, kern)))
    print("%s %s: %d instructions, reads even %d odd %d -> %d modelled cycles per iteration"
          % (cubin, kern, n, reads[0], reads[1], max(n, *reads)))
    for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:10]:
        print("   %-8s %d register reads: %d" % (k[0], k[1], v))
