#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 200 ./tools/exp_bdg 0 > gpurun_out/${tag}_exp_bdg.log 2>&1
timeout 200 ./tools/exp_bdg_o05 0 >> gpurun_out/${tag}_exp_bdg.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
timeout 900 python tools/bench_shapes.py 100 6 > gpurun_out/${tag}_shapes.txt 2>&1
timeout 300 python tools/bench_misc2.py > gpurun_out/${tag}_misc2.log 2>&1
cat gpurun_out/${tag}_exp_bdg.log; tail -6 gpurun_out/${tag}_tests.log; head -8 gpurun_out/${tag}_misc2.log; cat gpurun_out/${tag}_shapes.txt
