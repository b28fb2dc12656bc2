#!/bin/bash
# round-2 GPU session 21: the driver's round-end sequence on the final code -- whole GPU suite, smoke, default bench (with the reference arms)
set -u
mkdir -p gpurun_out
timeout -k 10 720 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider --durations=8 > gpurun_out/s21_tests.log 2>&1; echo "suite rc=$?"; tail -14 gpurun_out/s21_tests.log
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s21_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s21_smoke.log
timeout -k 10 600 python bench.py --breakdown gpurun_out/breakdown21_pose512.txt > gpurun_out/bench21_pose512.json 2> gpurun_out/bench21_pose512.err; echo "bench rc=$? $(head -c 700 gpurun_out/bench21_pose512.json)"; tail -2 gpurun_out/bench21_pose512.err
