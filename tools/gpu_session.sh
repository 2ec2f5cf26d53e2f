#!/bin/bash
# One GPU-box session: tests, experiments, bench.  Everything lands in gpurun_out/<tag>_*.log
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/${tag}_tests.log
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 300 ./tools/exp_bdg 4 > gpurun_out/${tag}_exp_bdg4.log 2>&1
timeout 300 python tools/bench_earlystop.py > gpurun_out/${tag}_earlystop.log 2>&1
timeout 300 python tools/bench_ops.py > gpurun_out/${tag}_ops.log 2>&1
PB_TRANSPOSE_NO_TMA=1 timeout 300 python tools/bench_ops.py > gpurun_out/${tag}_ops_notma.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_exp_bdg4.log; cat gpurun_out/${tag}_earlystop.log; tail -12 gpurun_out/${tag}_ops.log; head -c 3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
