#!/bin/bash
# N-GPU session: bench.py weak (overlapped and serial gather), strong
n=${2:-2}
tag=${1:-r02h}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:3}" > gpurun_out/${tag}_$2.json 2> gpurun_out/${tag}_$2.err; echo "rc=$?" >> gpurun_out/${tag}_$2.err; }
run 29521 weak --steps 4 --warmup 3 --no-e2e
run 29522 weak_serial --steps 4 --warmup 3 --no-e2e --no-extra --no-overlap
run 29523 strong --steps 4 --warmup 3 --no-e2e --no-extra --scaling strong
for f in weak weak_serial strong; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["timing"], {k:(v["value"],v["ms_per_step"],v.get("gather_ms"),v["roofline"].get("frac")) for k,v in d["extra"].items()})
except Exception as e:
    print("$f failed", e)
PY
tail -2 gpurun_out/${tag}_$f.err; done
