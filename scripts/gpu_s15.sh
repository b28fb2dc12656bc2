#!/bin/bash
# round-2 GPU session 15: thin-output kernels for the small heads, source-resolution SPADE norm reduction, tiled NCHW->NHWC pack,
# weight-gradient lanes: tests, then A/B of each switch (quick graph-replay timings), then a timeline with everything on
set -u
mkdir -p gpurun_out
FSV_THIN_OUT_MIN_PX=64 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/s15_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/s15_ops.log
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu --timeout 300 -p no:cacheprovider -k "spade or up2" > gpurun_out/s15_tc.log 2>&1; echo "tc rc=$?"; tail -3 gpurun_out/s15_tc.log
FSV_SIDE_LANES=4 FSV_THIN_OUT_MIN_PX=64 timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_nets.py tests/test_gpu_model.py -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/s15_lanes.log 2>&1; echo "lanes rc=$?"; tail -3 gpurun_out/s15_lanes.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 10 > gpurun_out/bench15_$name.json 2> gpurun_out/bench15_$name.err; echo "$name rc=$? $(head -c 120 gpurun_out/bench15_$name.json)"; }
run old FSV_PACK_TILED=0 FSV_SPADE_NORM_SRC=0
run pack FSV_PACK_TILED=1 FSV_SPADE_NORM_SRC=0
run base FSV_X=1
run thin FSV_THIN_OUT_MIN_PX=64
run lanes2 FSV_SIDE_LANES=2
run lanes4 FSV_SIDE_LANES=4
run all4 FSV_SIDE_LANES=4 FSV_THIN_OUT_MIN_PX=64
run all3 FSV_SIDE_LANES=3 FSV_THIN_OUT_MIN_PX=64
run base2 FSV_X=1
FSV_SIDE_LANES=4 FSV_THIN_OUT_MIN_PX=64 timeout 300 python scripts/trace_step.py --out gpurun_out/trace15_all > gpurun_out/s15_trace.log 2>&1; echo "trace rc=$?"; head -3 gpurun_out/trace15_all.txt
