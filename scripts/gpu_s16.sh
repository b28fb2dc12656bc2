#!/bin/bash
# round-2 GPU session 16: persistent fused-SPADE kernels (FSV_SPADE_PERSIST: 1 fwd, 2 bwd, 3 both), 128-pixel pack tiles, 4-pixel thin-input
# forward, float4 upsample: tests, A/B, timeline
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/s16_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/s16_ops.log
FSV_SPADE_PERSIST=1 timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -x -q -m gpu --timeout 120 -p no:cacheprovider -k "spade" > gpurun_out/s16_tc_p1.log 2>&1; echo "spade persist fwd rc=$?"; tail -3 gpurun_out/s16_tc_p1.log
FSV_SPADE_PERSIST=3 timeout -k 10 300 python -m pytest tests/test_gpu_tc.py -x -q -m gpu --timeout 120 -p no:cacheprovider -k "spade" > gpurun_out/s16_tc_p3.log 2>&1; echo "spade persist fwd+bwd rc=$?"; tail -3 gpurun_out/s16_tc_p3.log
FSV_SPADE_PERSIST=3 timeout -k 10 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_graph.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/s16_nets_p3.log 2>&1; echo "nets persist rc=$?"; tail -3 gpurun_out/s16_nets_p3.log
run() { name=$1; shift; env "$@" timeout -k 10 300 python bench.py --quick --steps 10 > gpurun_out/bench16_$name.json 2> gpurun_out/bench16_$name.err; echo "$name rc=$? $(head -c 120 gpurun_out/bench16_$name.json)"; }
run base FSV_X=1
run packold FSV_PACK_TILED=0
run sp1 FSV_SPADE_PERSIST=1
run sp2 FSV_SPADE_PERSIST=2
run sp3 FSV_SPADE_PERSIST=3
run base2 FSV_X=1
FSV_SPADE_PERSIST=3 timeout -k 10 300 python scripts/trace_step.py --out gpurun_out/trace16_sp3 > gpurun_out/s16_trace.log 2>&1; echo "trace rc=$?"; head -3 gpurun_out/trace16_sp3.txt
