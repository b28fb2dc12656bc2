#!/bin/bash
# round-2 GPU session 18: persistent SPADE with 16-byte bias loads and the transposed (whole-pixel) output stores: tests, A/B against the
# non-persistent kernels, timeline
set -u
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -x -q -m gpu --timeout 120 -p no:cacheprovider -k "spade" > gpurun_out/s18_spade.log 2>&1; echo "spade rc=$?"; tail -3 gpurun_out/s18_spade.log
timeout -k 10 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_graph.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/s18_nets.log 2>&1; echo "nets rc=$?"; tail -3 gpurun_out/s18_nets.log
run() { name=$1; shift; env "$@" timeout -k 10 300 python bench.py --quick --steps 10 > gpurun_out/bench18_$name.json 2> gpurun_out/bench18_$name.err; echo "$name rc=$? $(head -c 120 gpurun_out/bench18_$name.json)"; }
run base FSV_X=1
run sp0 FSV_SPADE_PERSIST=0
run sp2 FSV_SPADE_PERSIST=2
run base2 FSV_X=1
timeout -k 10 300 python scripts/trace_step.py --out gpurun_out/trace18 > gpurun_out/s18_trace.log 2>&1; echo "trace rc=$?"; head -3 gpurun_out/trace18.txt
