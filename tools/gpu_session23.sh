#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
PB_SHAPES_TS=2600,3000,3500,3840,4000,4096 PB_SHAPES_TR=1.0,0.72,0.5 python tools/bench_shapes.py 100 4 > gpurun_out/${tag}_shapes_long.txt 2>&1
cat gpurun_out/${tag}_shapes_long.txt; tail -3 gpurun_out/${tag}_tests.log
