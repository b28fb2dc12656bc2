#!/bin/bash
# round-2 GPU session 7 (N GPUs of one box): weak-scaling bench of the headline workload at N = 8, 4, 2, 1
set -u
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err
  echo "N=$n rc=$? $(head -c 330 gpurun_out/bench_scale_n$n.json)"; tail -1 gpurun_out/bench_scale_n$n.err | cut -c1-200
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err; echo "N=1 $(head -c 330 gpurun_out/bench_scale_n1.json)"
