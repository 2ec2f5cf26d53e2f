#!/bin/bash
tag=${1:-r02p}
n=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 tools/bench_gather.py > gpurun_out/${tag}_gather.log 2>&1
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/${tag}_gather.log | tail -5
