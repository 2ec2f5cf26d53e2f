#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
PB_SHAPES_TS=64,96,100,128,150,190 PB_SHAPES_TR=0.72 python tools/bench_shapes.py 100 6 > gpurun_out/${tag}_shapes.txt 2>&1
python tools/bench_configs.py cfg5 > gpurun_out/${tag}_cfg5.log 2>&1 || python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench.json 2>gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_shapes.txt; tail -3 gpurun_out/${tag}_cfg5.log
python - <<PY
import json, os
f="gpurun_out/${tag}_bench.json"
if os.path.exists(f):
    d=json.load(open(f))
    for k,v in d["extra"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline",{}).get("frac"))
PY
