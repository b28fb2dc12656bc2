#!/bin/bash
# usage: scripts/gpurun_retry_n.sh <gpus> <timeout_s> <logfile> <command...>
n=$1; t=$2; log=$3; shift 3
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$t" -- "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
