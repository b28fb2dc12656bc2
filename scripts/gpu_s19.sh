#!/bin/bash
# round-2 GPU session 19: shifted-descriptor experiments (halo-tile operand reuse feasibility)
set -u
mkdir -p gpurun_out
for m in 0 1 2; do FSV_TC_PERSIST=0 FSV_TC_SHIFT_EXP=$m timeout -k 10 120 python scripts/shift_exp.py 2>&1 | grep "FSV_TC_SHIFT_EXP=" ; done | tee gpurun_out/s19_shift_conv.log
for m in 0 1 2 3; do FSV_WG_SHIFT_EXP=$m timeout -k 10 120 python scripts/shift_exp.py 2>&1 | grep "FSV_WG_SHIFT_EXP=" ; done | tee gpurun_out/s19_shift_wg.log
