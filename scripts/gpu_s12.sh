#!/bin/bash
# round-2 GPU session 12: merged parity-class launches (stride-2 dgrad, up2 forward) + deep ring A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py tests/test_gpu_variants.py -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/s12_tests.log 2>&1; echo "tc tests rc=$?"; tail -5 gpurun_out/s12_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-baselines > gpurun_out/bench12_$name.json 2> gpurun_out/bench12_$name.err; echo "$name rc=$? $(head -c 260 gpurun_out/bench12_$name.json)"; }
run merged FSV_TC_MERGE=1
run unmerged FSV_TC_MERGE=0
run deep FSV_TC_DEEP=1
run merged2 FSV_TC_MERGE=1
