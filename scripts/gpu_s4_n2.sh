#!/bin/bash
# round-2 GPU session 4 (2 GPUs): data-parallel bench of the headline workload, both gradient-sync modes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_nets.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s4_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_n2_$name.json 2> gpurun_out/bench_n2_$name.err
  echo "$name rc=$? $(head -c 700 gpurun_out/bench_n2_$name.json)"; tail -2 gpurun_out/bench_n2_$name.err
}
run collect FSV_WGRAD_SIDE=1
run overlap FSV_WGRAD_SIDE=0
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-baselines > gpurun_out/bench_n1_ref.json 2> gpurun_out/bench_n1_ref.err; head -c 400 gpurun_out/bench_n1_ref.json; echo
timeout 600 python scripts/infer_sweep.py --sizes 256,512 --shots 1,5 --frames 8 --out gpurun_out/infer_sweep.jsonl
