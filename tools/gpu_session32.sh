#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python tools/bench_earlystop.py > gpurun_out/${tag}_earlystop.log 2>&1
PB_SHAPES_TS=200,240,256,300,320,350 PB_SHAPES_TR=1.0,0.72 python tools/bench_shapes.py 100 6 > gpurun_out/${tag}_shapes.txt 2>&1
tail -2 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_earlystop.log | tail -5; cat gpurun_out/${tag}_shapes.txt
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step_rank0"])
for k,v in d["extra"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("ms_each_call"), v.get("roofline",{}).get("frac"))
PY
