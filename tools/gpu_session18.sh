#!/bin/bash
tag=${1:-r02u}
mkdir -p gpurun_out
timeout 300 python tools/bench_earlystop.py > gpurun_out/${tag}_earlystop.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_earlystop.log; tail -6 gpurun_out/${tag}_tests.log
