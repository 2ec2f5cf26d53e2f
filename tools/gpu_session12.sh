#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python tools/bench_earlystop.py > gpurun_out/${tag}_earlystop.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "deconv or early" 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
timeout 600 python tools/debug_case.py > gpurun_out/${tag}_debug.log 2>&1
cat gpurun_out/${tag}_earlystop.log; tail -5 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_debug.log
