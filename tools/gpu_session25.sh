#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?"
python tools/bench_shapes.py 100 6 > gpurun_out/${tag}_shapes.txt 2>&1
tail -3 gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_smoke.log; cut -c1-600 gpurun_out/${tag}_bench.json; cut -c1-300 gpurun_out/${tag}_bench_ref.json; tail -3 gpurun_out/${tag}_shapes.txt
