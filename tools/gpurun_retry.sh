#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
# retries while gpurun answers "busy / no slot" (exit code 3 or a transient status), up to ~40 minutes
for attempt in $(seq 1 60); do
    out=$(/usr/local/graft/bin/gpurun --timeout "$1" "${@:2}" 2>&1)
    rc=$?
    echo "$out" | tail -80
    if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then
        sleep 45
        continue
    fi
    exit $rc
done
exit 3
