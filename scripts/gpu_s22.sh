#!/bin/bash
# round-2 GPU session 22: per-weight spectral tickets in reduction lanes (the session-21 failures were a ticket race between branch streams):
# whole GPU suite twice in fresh processes
set -u
mkdir -p gpurun_out
timeout -k 10 260 python -m pytest tests -q -m gpu --timeout 200 -p no:cacheprovider > gpurun_out/s22_tests_a.log 2>&1; echo "suite a rc=$?"; tail -4 gpurun_out/s22_tests_a.log | cut -c1-200
timeout -k 10 260 python -m pytest tests -q -m gpu --timeout 200 -p no:cacheprovider > gpurun_out/s22_tests_b.log 2>&1; echo "suite b rc=$?"; tail -4 gpurun_out/s22_tests_b.log | cut -c1-200
