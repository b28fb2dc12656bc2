#!/bin/bash
# round-2 GPU session 14: source-resolution backward of the upsample-collapsed convs (tests + A/B), kernel timeline of the graph replay,
# per-shape breakdown of the current code
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu --timeout 300 -p no:cacheprovider -k "up2 or wgrad" > gpurun_out/s14_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s14_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 10 > gpurun_out/bench14_$name.json 2> gpurun_out/bench14_$name.err; echo "$name rc=$? $(head -c 200 gpurun_out/bench14_$name.json)"; }
run base FSV_UP2_BWD=0
run up2src FSV_UP2_BWD=1
run base2 FSV_UP2_BWD=0
run up2src2 FSV_UP2_BWD=1
timeout 300 python scripts/trace_step.py --out gpurun_out/trace_graph > gpurun_out/s14_trace.log 2>&1; echo "trace rc=$?"; tail -8 gpurun_out/s14_trace.log
FSV_UP2_BWD=1 timeout 300 python bench.py --no-baselines --breakdown gpurun_out/breakdown14_up2src.txt > gpurun_out/bench14_full_up2src.json 2> gpurun_out/bench14_full_up2src.err; echo "full rc=$? $(head -c 200 gpurun_out/bench14_full_up2src.json)"
