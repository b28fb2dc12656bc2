#!/bin/bash
# round-2 GPU session 5: graph-vs-eager test, fused-loss tests, D-step stream / stream-priority A/B, headline bench, inference sweep
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_graph.py tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_variants.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/s5_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/s5_tests.log
echo "== A/B (quick, not bench values)"
for envs in "FSV_DSTEP_STREAM=0 FSV_MAIN_PRIORITY=0 FSV_GROUP_SPECTRAL_BWD=0" "FSV_DSTEP_STREAM=0 FSV_MAIN_PRIORITY=0 FSV_GROUP_SPECTRAL_BWD=1" "FSV_DSTEP_STREAM=0 FSV_MAIN_PRIORITY=-1" "FSV_DSTEP_STREAM=1 FSV_MAIN_PRIORITY=-1" "FSV_DSTEP_STREAM=1 FSV_MAIN_PRIORITY=0"; do
  for wl in pose512 face256; do
    echo "$envs $wl: $(env $envs timeout 300 python bench.py --quick --workload $wl --steps 10 --warmup 3 2>&1 | tail -1)"
  done
done
for wl in pose512 face256 street256x512 face256t pose512x256; do
  FSV_DSTEP_STREAM=1 timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-baselines > gpurun_out/bench5_dstep_$wl.json 2> gpurun_out/bench5_dstep_$wl.err; echo "bench dstep $wl rc=$? $(head -c 330 gpurun_out/bench5_dstep_$wl.json)"
done
timeout 600 python scripts/infer_sweep.py --sizes 256,512,1024 --shots 1,5 --frames 8 --out gpurun_out/infer_sweep.jsonl
