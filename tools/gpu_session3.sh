#!/bin/bash
# 2-GPU session: NCCL gather paths of bench.py (weak and strong), ES timing
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 300 python tools/bench_earlystop.py > gpurun_out/${tag}_earlystop.log 2>&1
timeout 200 python -m pytest tests -m gpu -q -x -k "early_stopping" 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "rc=$?" >> gpurun_out/${tag}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong --no-extra --no-e2e > gpurun_out/${tag}_bench_n2_strong.json 2> gpurun_out/${tag}_bench_n2_strong.err
echo "rc=$?" >> gpurun_out/${tag}_bench_n2_strong.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/${tag}_bench_n2_ref.json 2> gpurun_out/${tag}_bench_n2_ref.err
cat gpurun_out/${tag}_earlystop.log gpurun_out/${tag}_tests.log; head -c 2500 gpurun_out/${tag}_bench_n2.json; tail -3 gpurun_out/${tag}_bench_n2.err; head -c 1500 gpurun_out/${tag}_bench_n2_strong.json; tail -3 gpurun_out/${tag}_bench_n2_strong.err; head -c 600 gpurun_out/${tag}_bench_n2_ref.json
