#!/bin/bash
# round-2 GPU session 10 (2 GPUs): the NCCL gradient-sync check alone, bounded
set -u
mkdir -p gpurun_out
CHECK_DP_DUMP_S=90 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_dp.py > gpurun_out/dp.out 2> gpurun_out/dp.err
echo "check_dp rc=$?"; tail -3 gpurun_out/dp.out; grep -v Warning gpurun_out/dp.err | tail -60
