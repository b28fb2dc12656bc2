#!/usr/bin/env python
"""Round-2 ncu outputs in gpurun_out/ -> committed summaries under profiles/ (run in the build container)."""
import collections
import csv
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic']


def launches(src='gpurun_out/launches_r2.csv', take=2960):
    raw = [l for l in open(src) if not l.startswith('==')]
    r = csv.reader(raw)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    rows = [row for row in r if len(row) > vi]
    names = [re.sub(r'\(.*', '', row[ki])[:80] for row in rows]
    step, nm = rows[-take:], names[-take:]
    agg = collections.defaultdict(lambda: [0.0, 0])
    tot = 0.0
    for row, name in zip(step, nm):
        v = float(row[vi].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row[ui], 1)
        agg[name][0] += v
        agg[name][1] += 1
        tot += v
    out = ['ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 9000 python bench.py --quick --no-graph --steps 1 --warmup 1   (pose512, B=2)',
           'the last %d launches of the run ~ one eager training iteration (D-step + G-step incl. both Adam updates); cold-cache and' % take,
           'serialised under ncu: compare SHARES, not absolutes.  In the replayed CUDA graph the weight-gradient kernels (k_wgrad_tc_mn,',
           'k_conv_wgrad*, k_colsum, k_sn_dot / k_sn_bwd) run on a side stream and the reference-encoder branch on a third one.',
           'launches %d, total %.3f ms' % (take, tot / 1e6)]
    ours = sum(v[0] for k, v in agg.items() if k.startswith(('k_', 'void k_')))
    out.append("share of the time in this repo's kernels (k_*): %.1f%%; the rest: torch elementwise glue of the losses, allocation fills, cuBLAS for the" % (100 * ours / tot))
    out.append('  pre-summed up2 weights (6 tiny sgemm) -- see DESIGN.md section 7')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        out.append('%9.3f ms %5.1f%%  x%-5d %s' % (v[0] / 1e6, 100 * v[0] / tot, v[1], k))
    open('profiles/launches_r2_summary.txt', 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:12]))


def full():
    out = ['ncu --set full --clock-control none --import-source on (B200, round 2, pose512 eager step); .ncu-rep files stay in gpurun_out/ (scratch).']
    for rep in ('prof_conv_tc_r2', 'prof_spade_tc_r2', 'prof_wgrad_tc_r2'):
        path = 'gpurun_out/%s.ncu-rep' % rep
        if not os.path.exists(path):
            continue
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        r = list(csv.reader(txt.splitlines()))
        hdr = r[0]
        idx = [hdr.index(w) for w in WANT if w in hdr]
        out.append('== ' + rep + '.ncu-rep')
        for row in r[2:]:
            out.append('  ' + ' | '.join('%s=%s' % (hdr[i], row[i][:70]) for i in idx))
    open('profiles/ncu_full_r2_summary.txt', 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    launches()
    full()
