#!/bin/bash
tag=${1:-r02x}
mkdir -p gpurun_out
timeout 1500 python tools/sanitize_case.py > gpurun_out/${tag}_plain.log 2>&1; echo "plain rc=$?"
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/${tag}_plain.log; grep -E "ERROR SUMMARY|Invalid|done" gpurun_out/${tag}_memcheck.log | head -10
