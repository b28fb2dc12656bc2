#!/bin/bash
# round-2 GPU session 23 (2 GPUs): the NCCL gradient-sync check on the final code (weight-gradient lanes, ticket lanes), bounded
set -u
mkdir -p gpurun_out
CHECK_DP_DUMP_S=70 timeout -k 5 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_dp.py > gpurun_out/dp23.out 2> gpurun_out/dp23.err
echo "check_dp rc=$?"; tail -3 gpurun_out/dp23.out; grep -v Warning gpurun_out/dp23.err | tail -25
