#!/bin/bash
# ring-depth sweep of the tcgen05 conv / wgrad kernels (quick graph-replay timings; not bench values)
for cfg in "4 4" "0 4" "6 4" "4 0" "4 6" "0 0" "6 6"; do
    set -- $cfg
    tc=$1; wg=$2
    envs=""
    [ "$tc" != "0" ] && envs="$envs FSV_TC_STAGES=$tc"
    [ "$wg" != "0" ] && envs="$envs FSV_WG_STAGES=$wg"
    out=$(env $envs timeout 200 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1)
    echo "TC=$tc WG=$wg $out"
done
