#!/bin/bash
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 300 ./tools/exp_bdg 6 > gpurun_out/${tag}_exp_bdg6.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
timeout 900 python tools/bench_shapes.py 100 6 > gpurun_out/${tag}_shapes.txt 2>&1
cat gpurun_out/${tag}_exp_bdg6.log; tail -6 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_shapes.txt
