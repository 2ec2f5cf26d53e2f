#!/bin/bash
tag=${1:-r02w}
mkdir -p gpurun_out
timeout 300 python tools/debug_cfg2.py > gpurun_out/${tag}_cfg2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
head -30 gpurun_out/${tag}_cfg2.log
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"])
print(d["extra"]["cfg2_deconv_10k_x_300"]["ms_per_step"])
PY
