#!/bin/bash
# round-2 GPU session 3: full GPU suite, branch-stream / own-Adam A/B, headline bench, smoke, ncu launch list + full captures, inference sweep
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/s3_tests.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/s3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s3_smoke.log
echo "== A/B (quick, not bench values)"
for envs in "FSV_BRANCH_STREAMS=0" "FSV_BRANCH_STREAMS=1" "FSV_BRANCH_STREAMS=1 FSV_OWN_ADAM=0" "FSV_BRANCH_STREAMS=1 FSV_TC_PERSIST=0"; do
  for wl in pose512 face256; do
    echo "$envs $wl: $(env $envs timeout 300 python bench.py --quick --workload $wl --steps 10 --warmup 3 2>&1 | tail -1)"
  done
done
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/bd3_pose512.txt > gpurun_out/bench3_pose512.json 2> gpurun_out/bench3_pose512.err; echo "bench rc=$?"; head -c 1000 gpurun_out/bench3_pose512.json; echo; tail -2 gpurun_out/bench3_pose512.err
# ncu: launch list of one eager step (cold-cache, serialised: shares only), then full captures of the dominant kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 9000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu launch list rc=$?"; tail -1 gpurun_out/ncu_launch_r2.log
for spec in "k_conv_tc_p:4:3:prof_conv_tc_r2" "k_spade_tc:14:4:prof_spade_tc_r2" "k_wgrad_tc_mn:30:2:prof_wgrad_tc_r2"; do
    IFS=: read -r kern skip cnt out <<< "$spec"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c $cnt -o gpurun_out/$out -f \
        python bench.py --quick --no-graph --steps 1 --warmup 0 > gpurun_out/ncu_$out.log 2>&1
    echo "ncu $out rc=$?"
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
timeout 900 python scripts/infer_sweep.py --sizes 256,512 --shots 1,5 --frames 8 --out gpurun_out/infer_sweep.jsonl
