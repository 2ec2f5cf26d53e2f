#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
CMD="python bench.py --voxels 8880 --scans 1200 --t-r 0.72 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fast_bdc -s 3 -c 1 -f -o gpurun_out/${tag}_bd_t1200 $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_plain.log | cut -c1-400; tail -3 gpurun_out/${tag}_ncu.log
