"""Per-kernel summary of a binary / object: registers, spills, hot-loop instruction mix and the static
register-file / issue model of tools/rf_model.py.

    python tools/kstat.py tools/exp_bdg [kernel-name-substring]
"""
import collections
import re
import subprocess
import sys

import rf_model as m


def kernels(path):
    out = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    res = {}
    name = None
    for ln in out.splitlines():
        mm = re.search(r"Function (\S+):", ln)
        if mm:
            name = mm.group(1)
        mm = re.search(r"REG:(\d+).*?STACK:(\d+).*?SHARED:(\d+)", ln)
        if mm and name:
            res[name] = (int(mm.group(1)), int(mm.group(2)))
    return res


if __name__ == "__main__":
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else "fast_bdg_kernel"
    allsass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
    for name, (regs, stack) in kernels(path).items():
        if sub not in name:
            continue
        on, lines = False, []
        for ln in allsass:
            if "Function :" in ln:
                on = name in ln
            elif on and re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", ln):
                lines.append(ln)
        loop = m.hot_loop(lines)
        n, reads, per = m.model(loop)
        c = collections.Counter()
        for ln in loop:
            mm = re.search(r"\*/\s+(@!?U?P\d\s+)?(\S+)\s*(.*?);", ln)
            if mm:
                c[mm.group(2).split(".")[0]] += 1
        tag = re.sub(r"^_ZN2pb\d+", "", name)[:70]
        print("%-72s regs %3d stack %3d | loop %4d instr, reads %4d/%4d -> %4d cycles | FFMA3 %3d | %s"
              % (tag, regs, stack, n, reads[0], reads[1], max(n, *reads), per.get(("FFMA", 3), 0),
                 " ".join("%s:%d" % kv for kv in c.most_common(9))))
