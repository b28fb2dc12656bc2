#!/bin/bash
# round-2 GPU session 13: whole GPU suite with -x on the current code (out-of-line transcendental activations, leaner norm reductions,
# merged parity classes), default bench twice
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/s13_tests.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/s13_tests.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-baselines > gpurun_out/bench13_$name.json 2> gpurun_out/bench13_$name.err; echo "$name rc=$? $(head -c 260 gpurun_out/bench13_$name.json)"; }
run a FSV_X=1
run b FSV_X=1
