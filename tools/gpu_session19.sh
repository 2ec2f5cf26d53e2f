#!/bin/bash
tag=${1:-r02v}
mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err ) 2> gpurun_out/${tag}_bench_ref.time
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2> gpurun_out/${tag}_bench.time
python - <<PY
import json
r=json.loads(open("gpurun_out/${tag}_bench_ref.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["cpu_baseline"])
print("ours", d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["roofline"]["frac_of_microbench"], d["roofline"]["traffic"], d["cpu_baseline"], d["clocks"])
print("same config:", r["config"] == d["config"], "ratio e2e", d["e2e"]["value"]/r["value"])
for k,v in d["extra"].items(): print(k, v["value"], v["unit"], v["ms_per_step"], v["roofline"].get("frac"), v["roofline"].get("kernel_ms"))
PY
cat gpurun_out/${tag}_bench_ref.time gpurun_out/${tag}_bench.time | grep real
