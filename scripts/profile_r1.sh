#!/bin/bash
# Round-1 profiling recipe (run under gpurun on one B200); outputs go to gpurun_out/, summaries are copied to profiles/.
set -u
mkdir -p gpurun_out
# 1. launch list of ~1.5 training steps (eager launches so that every kernel is a separate ncu record)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 6000 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log
# 2. full captures of the dominant kernels (first launches of the generator forward = the 256x256 layers)
for spec in "k_conv_tc:2:3:prof_conv_tc_r1" "k_spade_tc:20:2:prof_spade_tc_r1" "k_wgrad_tc_mn:30:2:prof_wgrad_tc_r1"; do
    IFS=: read -r kern skip cnt out <<< "$spec"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c $cnt -o gpurun_out/$out -f \
        python bench.py --quick --no-graph --steps 1 --warmup 0 > gpurun_out/ncu_$out.log 2>&1
    tail -1 gpurun_out/ncu_$out.log
done
ls -la gpurun_out/*.ncu-rep
