#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
python tools/bench_ops.py > gpurun_out/${tag}_ops.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rows_conv8_kernel|transpose_tma_kernel" -c 6 -f -o gpurun_out/${tag}_ops python tools/bench_ops.py > gpurun_out/${tag}_ops_ncu.log 2>&1
cat gpurun_out/${tag}_ops.log; tail -3 gpurun_out/${tag}_ops_ncu.log
