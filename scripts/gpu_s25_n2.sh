#!/bin/bash
# round-2 GPU session 25 (2 GPUs): one bench line of the final code at N = 2 (weak scaling), bounded
set -u
mkdir -p gpurun_out
timeout -k 5 55 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench25_n2.json 2> gpurun_out/bench25_n2.err
echo "n2 rc=$? $(head -c 300 gpurun_out/bench25_n2.json)"; tail -2 gpurun_out/bench25_n2.err | cut -c1-300
