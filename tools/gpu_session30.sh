#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"
{ uptime; ps -eo pid,ppid,pcpu,etime,comm --sort=-pcpu | head -12; } > gpurun_out/${tag}_ps.log 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
{ uptime; ps -eo pid,ppid,pcpu,etime,comm --sort=-pcpu | head -8; } >> gpurun_out/${tag}_ps.log 2>&1
cat gpurun_out/${tag}_ps.log
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print(d["value"], d["e2e"]["ms_per_step_rank0"])
for k,v in d["extra"].items(): print(k, v.get("ms_per_step"), v.get("ms_each_call"), v.get("roofline",{}).get("frac"))
PY
