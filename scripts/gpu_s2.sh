#!/bin/bash
# round-2 GPU session 2: new kernels' tests, pose512 / face256 bench with breakdowns, A/B of grouped spectral norm / side-stream wgrad / persistent conv
set -u
mkdir -p gpurun_out
export FSV_WGRAD_SIDE=0
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_nets.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s2_tests.log 2>&1; echo "tests (side stream off) rc=$?"; tail -5 gpurun_out/s2_tests.log
FSV_WGRAD_SIDE=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_nets.py tests/test_gpu_tc.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s2_tests_side.log 2>&1; echo "tests (side stream on) rc=$?"; tail -5 gpurun_out/s2_tests_side.log
timeout 1200 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/s2_dropin.log 2>&1; echo "dropin rc=$?"; grep -E "drop-in|yardstick|passed|failed|^E  " gpurun_out/s2_dropin.log | tail -30
echo "== A/B (quick, not bench values)"
for envs in "FSV_GROUP_SPECTRAL=0 FSV_WGRAD_SIDE=0 FSV_TC_BN_OCC=0" "FSV_GROUP_SPECTRAL=1 FSV_WGRAD_SIDE=0 FSV_TC_BN_OCC=0" "FSV_GROUP_SPECTRAL=1 FSV_WGRAD_SIDE=1 FSV_TC_BN_OCC=0" "FSV_GROUP_SPECTRAL=1 FSV_WGRAD_SIDE=1 FSV_TC_BN_OCC=1" "FSV_GROUP_SPECTRAL=1 FSV_WGRAD_SIDE=1 FSV_TC_BN_OCC=1 FSV_TC_PERSIST=1"; do
  for wl in pose512 face256; do
    echo "$envs $wl: $(env $envs timeout 300 python bench.py --quick --workload $wl --steps 10 --warmup 3 2>&1 | tail -1)"
  done
done
export FSV_WGRAD_SIDE=1
for wl in pose512 face256; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --breakdown gpurun_out/bd_$wl.txt > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; head -c 1800 gpurun_out/bench_$wl.json; echo; tail -3 gpurun_out/bench_$wl.err
done
timeout 600 python bench.py --workload street256x512 --steps 10 --no-baselines > gpurun_out/bench_street.json 2> gpurun_out/bench_street.err; head -c 600 gpurun_out/bench_street.json; echo
timeout 600 python bench.py --workload face256t --steps 10 --no-baselines > gpurun_out/bench_face256t.json 2> gpurun_out/bench_face256t.err; head -c 600 gpurun_out/bench_face256t.json; echo
