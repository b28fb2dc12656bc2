#!/bin/bash
# round-2 GPU session 11: the driver's round-end sequence on the final code -- GPU suite with -x, smoke, default bench -- plus one A/B
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/s11_tests.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/s11_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s11_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s11_smoke.log
timeout 900 python bench.py > gpurun_out/bench11_pose512.json 2> gpurun_out/bench11_pose512.err; echo "bench rc=$? $(head -c 600 gpurun_out/bench11_pose512.json)"; tail -2 gpurun_out/bench11_pose512.err
FSV_NORM_BWD2=0 timeout 300 python bench.py --no-baselines > gpurun_out/bench11_nobwd2.json 2> gpurun_out/bench11_nobwd2.err; echo "nobwd2 rc=$? $(head -c 300 gpurun_out/bench11_nobwd2.json)"
timeout 300 python bench.py --no-baselines > gpurun_out/bench11_again.json 2> gpurun_out/bench11_again.err; echo "again rc=$? $(head -c 300 gpurun_out/bench11_again.json)"
