#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 300 ./tools/exp_bdg 7 > gpurun_out/${tag}_exp_bdg7.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_exp_bdg7.log; tail -10 gpurun_out/${tag}_tests.log
