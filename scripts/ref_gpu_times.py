#!/usr/bin/env python
"""Time the unmodified reference on the GPU (cuDNN/ATen) for a list of workloads: `name:kind:H:W:batch[:extra...]`.
One subprocess per (workload, tf32) so that cudnn.benchmark caches and allocator state do not leak between rows.
Output: one JSON line per row, appended to --out.  Profiling aid for BASELINE.md section 3 (bench.py --impl reference-gpu is
the contract arm)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(spec, tf32, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_arm
    f = spec.split(':')
    wl = dict(kind=f[1], H=int(f[2]), W=int(f[3]), ref_extra=f[5:])
    if '--temporal' in wl['ref_extra']:
        wl['ref_extra'].remove('--temporal')
        wl['temporal'] = True
    r = ref_arm.run_gpu(wl, int(f[4]), steps, warmup, tf32=tf32)
    r.update(name=f[0], spec=spec)
    print('RESULT ' + json.dumps(r), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('specs', nargs='*')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'ref_gpu_times.jsonl'))
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--child', default=None)
    ap.add_argument('--tf32', type=int, default=1)
    a = ap.parse_args()
    if a.child:
        child(a.child, bool(a.tf32), a.steps, a.warmup)
        sys.exit(0)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    for spec in a.specs:
        for tf32 in (1, 0):
            p = subprocess.run([sys.executable, __file__, '--child', spec, '--tf32', str(tf32), '--steps', str(a.steps), '--warmup', str(a.warmup)],
                               capture_output=True, text=True, timeout=900)
            line = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')]
            rec = json.loads(line[-1][7:]) if line else dict(spec=spec, tf32=bool(tf32), error=(p.stderr or p.stdout)[-1500:])
            with open(a.out, 'a') as f:
                f.write(json.dumps(rec) + '\n')
            print(json.dumps({k: v for k, v in rec.items() if k != 'losses'}), flush=True)
