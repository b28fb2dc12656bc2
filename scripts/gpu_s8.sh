#!/bin/bash
# round-2 GPU session 8: the whole GPU suite (no -x), VGG workload bench
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/s8_tests.log 2>&1; echo "suite rc=$?"; grep -E "graph vs eager|passed|failed|^FAILED" gpurun_out/s8_tests.log | tail -12
timeout 900 python bench.py --workload pose512vgg --steps 10 --warmup 3 > gpurun_out/bench8_pose512vgg.json 2> gpurun_out/bench8_pose512vgg.err; echo "bench pose512vgg rc=$? $(head -c 330 gpurun_out/bench8_pose512vgg.json)"; tail -2 gpurun_out/bench8_pose512vgg.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench8_pose512vgg.json'))
    print('vgg workload:', d['value'], 'frames/s; reference gpu', d.get('reference_gpu', {}).get('value'), d.get('vs_reference_gpu'))
except Exception as e:
    print('no vgg line', e)
PY
