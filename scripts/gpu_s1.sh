#!/bin/bash
# round-2 GPU session 1: existing suite + bring-up tests + live drop-in tests + reference-on-GPU timings
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
python -c "import os;print('cpus',os.cpu_count())"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_dropin.py > gpurun_out/s1_tests.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/s1_tests.log
FSV_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/s1_exp.log 2>&1; echo "experimental rc=$?"; tail -15 gpurun_out/s1_exp.log
timeout 1500 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/s1_dropin.log 2>&1; echo "dropin rc=$?"; grep -E "drop-in|passed|failed|Error|assert" gpurun_out/s1_dropin.log | tail -40
timeout 1500 python scripts/ref_gpu_times.py face256:face:256:256:8 pose512:pose:512:512:2 pose512x256:pose:512:256:2 street256x512:street:256:512:6 face256t:face:256:256:8:--temporal
