#!/usr/bin/env python
"""Turn the ncu outputs in gpurun_out/ into the committed summaries under profiles/ (run in the build container)."""
import collections, csv, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)

def launches(src='gpurun_out/launches_r1b.csv'):
    """ncu launch list of an eager run (init + warm-up step + one timed step): find the step period in the kernel-name
    sequence, keep the LAST full step (-> profiles/launches_r1.csv) and summarise it."""
    raw = [l for l in open(src) if not l.startswith('==')]
    r = csv.reader(raw); hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    rows = [row for row in r if len(row) > vi]
    names = [re.sub(r'\(.*', '', row[ki])[:80] for row in rows]
    n = len(names)
    period = next(p for p in range(1500, n // 2) if names[n - p:] == names[n - 2 * p:n - p])
    step = rows[n - period:]
    with open('profiles/launches_r1.csv', 'w', newline='') as f:
        w = csv.writer(f); w.writerow(hdr); w.writerows(step)
    agg = collections.defaultdict(lambda: [0.0, 0]); tot = 0.0
    for row, name in zip(step, names[n - period:]):
        v = float(row[vi].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row[ui], 1)
        agg[name][0] += v; agg[name][1] += 1; tot += v
    out = ['ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 python bench.py --quick --no-graph --steps 1 --warmup 1',
           '(the last full training step of the run = %d consecutive launches, found by the period of the kernel-name sequence;' % period,
           ' cold-cache, serialised: compare SHARES, not absolutes.  Taken at commit 51c2db7 (fused spectral norm); the later',
           ' thin-output / norm-reduction commits remove ~380 more launches, see DESIGN.md section 7)',
           'launches %d, total %.3f ms' % (period, tot / 1e6)]
    ours = sum(v[0] for k, v in agg.items() if k.startswith(('k_', 'void k_')))
    out.append('share of the time in this repo\'s kernels (k_*): %.1f%%; the rest is parameter-side torch work (Adam, losses, layout copies)' % (100 * ours / tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:50]:
        out.append('%9.3f ms %5.1f%%  x%-5d %s' % (v[0] / 1e6, 100 * v[0] / tot, v[1], k))
    open('profiles/launches_r1_summary.txt', 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:30]))

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']

def full():
    out = ['ncu --set full --clock-control none --import-source on (B200, round 1); .ncu-rep files stay in gpurun_out/ (scratch).']
    for k in ('conv_tc', 'spade_tc', 'wgrad_tc'):
        rep = 'gpurun_out/prof_%s_r1.ncu-rep' % k
        if not os.path.exists(rep): continue
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(w) for w in WANT if w in hdr]
        out.append('== ' + os.path.basename(rep))
        for row in rows[2:]:
            out.append('  ' + ' | '.join('%s=%s%s' % (hdr[i], row[i], (' ' + units[i]) if units[i] else '') for i in idx))
    return out

if __name__ == '__main__':
    launches()
    if '--launches-only' in sys.argv:
        sys.exit(0)
    o = full()
    print('\n'.join(o))
    open('profiles/ncu_full_r1_summary.txt', 'w').write('\n'.join(o) + '\n')
