#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
timeout 900 python tools/bench_shapes.py 50 > gpurun_out/${tag}_shapes.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
$CMD > gpurun_out/${tag}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fast_bdg -s 3 -c 1 -f -o gpurun_out/${tag}_bd_t300 $CMD > gpurun_out/${tag}_ncu2.log 2>&1
tail -6 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_shapes.txt; tail -3 gpurun_out/${tag}_ncu1.log; tail -3 gpurun_out/${tag}_ncu2.log; ls -la gpurun_out | tail -8
