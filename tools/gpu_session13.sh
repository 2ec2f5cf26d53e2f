#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 600 python tools/debug_case.py > gpurun_out/${tag}_debug.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "short_tr or long_series or variants" 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_debug.log; tail -5 gpurun_out/${tag}_tests.log
