#!/bin/bash
# final multi-GPU check: N = $2 ranks, both arms as the driver launches them
tag=${1:-r02x}; N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref_n$N.json 2> gpurun_out/${tag}_bench_ref_n$N.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/${tag}_bench_n$N.json; tail -3 gpurun_out/${tag}_bench_n$N.err; cut -c1-300 gpurun_out/${tag}_bench_ref_n$N.json
