#!/bin/bash
# 8-GPU run of the driver's command (weak scaling, default flags) + reference arm
tag=${1:-r02n}
n=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
echo "rc=$?" >> gpurun_out/${tag}_bench_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_n$n.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["timing"]["solver_ms"], d["timing"]["gather_ms"], d["timing"]["gather_exposed_ms"], d["e2e"], d["roofline"]["frac"])
for k,v in d["extra"].items(): print(k, v["value"], v["ms_per_step"], v.get("solver_ms"), v.get("gather_ms"), v["roofline"].get("frac"))
PY
tail -3 gpurun_out/${tag}_bench_n$n.err
