#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 600 python tools/bench_e2e.py > gpurun_out/${tag}_e2e.log 2>&1
OMP_NUM_THREADS=1 timeout 600 python tools/bench_e2e.py > gpurun_out/${tag}_e2e_omp1.log 2>&1
cat gpurun_out/${tag}_e2e.log; echo "--- OMP_NUM_THREADS=1"; cat gpurun_out/${tag}_e2e_omp1.log
