#!/bin/bash
# round-2 GPU session 9 (2 GPUs): gradient sync fired on the side stream during backward + chunked spectral groups
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s9_dp.log 2>&1; echo "dp test rc=$?"; tail -15 gpurun_out/s9_dp.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_n2_$name.json 2> gpurun_out/bench_n2_$name.err
  echo "$name rc=$? $(head -c 500 gpurun_out/bench_n2_$name.json)"; tail -2 gpurun_out/bench_n2_$name.err
}
run stream4 FSV_SYNC_STREAM=1
run stream1 FSV_SYNC_STREAM=1 FSV_SPECTRAL_CHUNKS=1
run collect FSV_SYNC_STREAM=0 FSV_SPECTRAL_CHUNKS=1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_n1_s9.json 2> gpurun_out/bench_n1_s9.err; head -c 400 gpurun_out/bench_n1_s9.json; echo
