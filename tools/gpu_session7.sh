#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 300 python tools/debug_case.py > gpurun_out/${tag}_debug.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${tag}_tests.log
timeout 900 python tools/bench_shapes.py 100 6 > gpurun_out/${tag}_shapes.txt 2>&1
cat gpurun_out/${tag}_debug.log; tail -10 gpurun_out/${tag}_tests.log; head -40 gpurun_out/${tag}_shapes.txt
