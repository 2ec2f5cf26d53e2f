#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
PB_SHAPES_TR=0.5,0.32 python tools/bench_shapes.py 100 4 > gpurun_out/${tag}_shapes_k40_k63.txt 2>&1
cat gpurun_out/${tag}_shapes_k40_k63.txt; tail -3 gpurun_out/${tag}_tests.log
