#!/bin/bash
# round-2 GPU session 6: SPADE channel-block A/B, full GPU suite on the final defaults, headline bench with baselines, other workloads
set -u
mkdir -p gpurun_out
echo "== A/B (quick, not bench values)"
for envs in "FSV_SPADE_CB=0" "FSV_SPADE_CB=32" "FSV_SPADE_CB=32 FSV_SPADE_CB_BWD=32" "FSV_SPADE_CB=0 FSV_SPADE_CB_BWD=32"; do
  for wl in pose512 face256; do
    echo "$envs $wl: $(env $envs timeout 300 python bench.py --quick --workload $wl --steps 10 --warmup 3 2>&1 | tail -1)"
  done
done
for envs in "FSV_TC_STAGES=4" "FSV_TC_STAGES=6" "FSV_WG_STAGES=6" "FSV_TC_STAGES=6 FSV_WG_STAGES=6"; do
  for wl in pose512; do
    echo "$envs $wl: $(env $envs timeout 300 python bench.py --quick --workload $wl --steps 10 --warmup 3 2>&1 | tail -1)"
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/s6_tests.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/s6_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s6_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s6_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/bd6_pose512.txt > gpurun_out/bench6_pose512.json 2> gpurun_out/bench6_pose512.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench6_pose512.json; echo; tail -2 gpurun_out/bench6_pose512.err
timeout 600 python scripts/infer_sweep.py --sizes 256,512,1024 --shots 1,5 --frames 8 --out gpurun_out/infer_sweep6.jsonl | cut -c1-400
timeout 300 python scripts/infer_sweep.py --sizes 256 --shots 20 --frames 6 --out gpurun_out/infer_sweep6.jsonl | cut -c1-400
timeout 900 python bench.py --workload pose512vgg --steps 10 --warmup 3 > gpurun_out/bench6_pose512vgg.json 2> gpurun_out/bench6_pose512vgg.err; echo "bench pose512vgg rc=$? $(head -c 330 gpurun_out/bench6_pose512vgg.json)"; tail -2 gpurun_out/bench6_pose512vgg.err
for wl in face256 street256x512 face256t pose512x256; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-baselines > gpurun_out/bench6_$wl.json 2> gpurun_out/bench6_$wl.err; echo "bench $wl rc=$? $(head -c 330 gpurun_out/bench6_$wl.json)"
done
