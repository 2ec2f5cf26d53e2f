#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
PB_SHAPES_TS=2600,3000,3072,3500,3840,4000,4096 PB_SHAPES_TR=1.0,0.72,0.5 python tools/bench_shapes.py 100 4 > gpurun_out/${tag}_shapes_long.txt 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_shapes_long.txt; tail -3 gpurun_out/${tag}_tests.log; cut -c1-300 gpurun_out/${tag}_bench.json
