#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"
for i in 1 2; do
python bench.py > gpurun_out/${tag}_bench$i.json 2> gpurun_out/${tag}_bench$i.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench$i.json"))
print(d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step_rank0"])
for k,v in d["extra"].items(): print(k, v.get("ms_per_step"), v.get("ms_each_call"), v.get("roofline",{}).get("frac"))
PY
done
tail -2 gpurun_out/${tag}_tests.log
