#!/bin/bash
tag=${1:-r02s}
mkdir -p gpurun_out
timeout 300 ./tools/exp_bdg 8 > gpurun_out/${tag}_exp_bdg8.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_exp_bdg8.log; tail -4 gpurun_out/${tag}_tests.log
