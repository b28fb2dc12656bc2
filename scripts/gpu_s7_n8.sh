#!/bin/bash
# round-2 GPU session 7 (8 GPUs of one box): weak-scaling bench of the headline workload at N = 8 and N = 1 on the same box
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_scale_n8.json 2> gpurun_out/bench_scale_n8.err
echo "N=8 rc=$? $(head -c 330 gpurun_out/bench_scale_n8.json)"; tail -2 gpurun_out/bench_scale_n8.err | cut -c1-300
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err; echo "N=1 $(head -c 330 gpurun_out/bench_scale_n1.json)"
