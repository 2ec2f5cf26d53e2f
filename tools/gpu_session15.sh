#!/bin/bash
tag=${1:-r02o}
n=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer_gather.py -q -x 2>&1 | tail -15 > gpurun_out/${tag}_peer_test.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:3}" > gpurun_out/${tag}_$2.json 2> gpurun_out/${tag}_$2.err; echo "rc=$?" >> gpurun_out/${tag}_$2.err; }
run 29541 peer --steps 4 --warmup 3 --no-e2e
run 29542 nccl_serial --steps 4 --warmup 3 --no-e2e --no-extra --no-overlap --gather-via nccl
cat gpurun_out/${tag}_peer_test.log
for f in peer nccl_serial; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["timing"], {k:(v["value"],v["ms_per_step"],v.get("solver_ms"),v.get("gather_ms"),v["roofline"].get("frac")) for k,v in d["extra"].items()})
except Exception as e:
    print("$f failed", e)
PY
tail -4 gpurun_out/${tag}_$f.err; done
