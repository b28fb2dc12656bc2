#!/bin/bash
# round-2 GPU session 17: ncu --set full on layers of known shape (scripts/ncu_layers.py) with the persistent SPADE kernels, ring-depth A/B
set -u
mkdir -p gpurun_out
export FSV_SPADE_PERSIST=3
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc_p|k_spade_tc|k_wgrad_tc_mn' -o gpurun_out/prof_layers_r2 -f \
    python scripts/ncu_layers.py gpurun_out/ncu_layers.json > gpurun_out/s17_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/s17_ncu.log
ls -la gpurun_out/*.ncu-rep
run() { name=$1; shift; env "$@" timeout -k 10 300 python bench.py --quick --steps 10 > gpurun_out/bench17_$name.json 2> gpurun_out/bench17_$name.err; echo "$name rc=$? $(head -c 120 gpurun_out/bench17_$name.json)"; }
run base FSV_X=1
run deep FSV_TC_DEEP=1
run st4 FSV_TC_STAGES=4
run wg6 FSV_WG_STAGES=6
run base2 FSV_X=1
