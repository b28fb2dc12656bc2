#!/bin/bash
# round-2 GPU session 20: reference encoders on separate streams (FSV_ENC_SPLIT): tests with the switch on, A/B; lane-count A/B
set -u
mkdir -p gpurun_out
FSV_ENC_SPLIT=1 timeout -k 10 500 python -m pytest tests/test_gpu_nets.py tests/test_gpu_graph.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/s20_split.log 2>&1; echo "enc split tests rc=$?"; tail -3 gpurun_out/s20_split.log
run() { name=$1; shift; env "$@" timeout -k 10 200 python bench.py --quick --steps 10 > gpurun_out/bench20_$name.json 2> gpurun_out/bench20_$name.err; echo "$name rc=$? $(head -c 120 gpurun_out/bench20_$name.json)"; }
run base FSV_X=1
run split FSV_ENC_SPLIT=1
run lanes5 FSV_SIDE_LANES=5
run split2 FSV_ENC_SPLIT=1
run base2 FSV_X=1
