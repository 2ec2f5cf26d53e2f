#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/${tag}_tests.log; tail -1 gpurun_out/${tag}_smoke.log; wc -l gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench_ref.json; cut -c1-200 gpurun_out/${tag}_bench.json
