#!/bin/bash
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "early_stopping or layout or row_kernels or deconv or ops" 2>&1 | tail -30 > gpurun_out/${tag}_tests.log
timeout 300 python tools/bench_earlystop.py > gpurun_out/${tag}_earlystop.log 2>&1
timeout 300 python tools/bench_ops.py > gpurun_out/${tag}_ops.log 2>&1
PB_ROWS_CONV4=1 PB_TRANSPOSE_NO_TMA=1 timeout 300 python tools/bench_ops.py > gpurun_out/${tag}_ops_r01.log 2>&1
timeout 300 python tools/bench_misc2.py > gpurun_out/${tag}_misc2.log 2>&1
tail -5 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_earlystop.log; tail -9 gpurun_out/${tag}_ops.log; tail -9 gpurun_out/${tag}_ops_r01.log; cat gpurun_out/${tag}_misc2.log
