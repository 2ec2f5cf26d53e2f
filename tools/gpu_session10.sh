#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${tag}_bench_e2e.json 2> gpurun_out/${tag}_bench_e2e.err
PB_NO_QUEUE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${tag}_bench_e2e_noq.json 2> gpurun_out/${tag}_bench_e2e_noq.err
tail -10 gpurun_out/${tag}_tests.log
for f in bench_e2e bench_e2e_noq; do python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_$f.json").read().strip().splitlines()[-1])
print("$f", d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"])
PY
done
