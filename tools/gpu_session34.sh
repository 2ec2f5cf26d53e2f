#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fast_bdg -s 3 -c 1 -f -o gpurun_out/${tag}_bd_t300 $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_plain.log | cut -c1-200; tail -3 gpurun_out/${tag}_ncu.log
