#!/bin/bash
# round-2 GPU session 24: one default-workload bench line of the very last code (after the ticket-lane fix), no reference sub-arms
set -u
mkdir -p gpurun_out
timeout -k 5 150 python bench.py --no-baselines > gpurun_out/bench24_pose512.json 2> gpurun_out/bench24_pose512.err; echo "bench rc=$? $(head -c 330 gpurun_out/bench24_pose512.json)"
