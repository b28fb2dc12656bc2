#!/usr/bin/env python
"""Per-kernel timeline of one training step (profiling aid, not a bench value).

    python scripts/trace_step.py [--workload pose512] [--out gpurun_out/trace] [--eager]

Builds the step exactly as bench.py does, replays the CUDA graph (or, with --eager, runs the Python-issued step) under
torch.profiler (CUPTI kernel activity records: start, duration, stream of every kernel) and writes
  <out>.jsonl  one line per kernel {name, ts_us, dur_us, stream} of ONE step, ordered by start time
  <out>.txt    per-stream busy time, the step's wall time, and the idle gaps of the busiest stream
nsys is not in the image; this is the closest thing to its timeline.  A number taken from this run is never a bench value.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts the package and baseline/ on sys.path)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='pose512')
    ap.add_argument('--out', default='gpurun_out/trace')
    ap.add_argument('--eager', action='store_true')
    args = ap.parse_args()
    from fsv import trainer, model
    dev = torch.device('cuda', 0)
    torch.cuda.set_stream(torch.cuda.Stream(priority=int(os.environ.get('FSV_MAIN_PRIORITY', '-1'))))
    wl = bench.WORKLOADS[args.workload]
    opt = bench.make_opt(args.workload)
    torch.manual_seed(0)
    step = model.Vid2VidStep(opt)
    if wl.get('temporal'):
        step.init_temporal_model()
    for m in [step.netG] + step.d_modules():
        m.train()
    optG, optD = trainer.make_step_optimizers(opt, step, capturable=not args.eager)
    devin = {k: v.to(dev) for k, v in bench.synth_inputs(args.workload, wl['batch'], seed=1234).items()}

    def eager():
        return trainer.train_iteration(step, optG, optD, devin)
    eager()
    run = eager if args.eager else trainer.GraphedStep(step, optG, optD, devin)
    for _ in range(3):
        run(devin) if not args.eager else run()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            run(devin) if not args.eager else run()
            torch.cuda.synchronize()
    # kernel records (with stream ids) are read back from the chrome trace
    tmp = args.out + '.chrome.json'
    prof.export_chrome_trace(tmp)
    tr = json.load(open(tmp))
    ks = [t for t in tr['traceEvents'] if t.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'ts' in t]
    ks.sort(key=lambda t: t['ts'])
    # split the two profiled steps at the largest gap near the middle
    mid = len(ks) // 2
    gaps = [(ks[i + 1]['ts'] - (ks[i]['ts'] + ks[i]['dur']), i) for i in range(max(1, mid - 200), min(len(ks) - 1, mid + 200))]
    cut = max(gaps)[1] + 1
    one = ks[cut:]
    t0 = one[0]['ts']
    with open(args.out + '.jsonl', 'w') as f:
        for t in one:
            f.write(json.dumps({'name': t['name'][:120], 'ts_us': round(t['ts'] - t0, 3), 'dur_us': round(t['dur'], 3),
                                'stream': t.get('args', {}).get('stream'), 'cat': t['cat'], 'grid': t.get('args', {}).get('grid'),
                                'block': t.get('args', {}).get('block'), 'smem': t.get('args', {}).get('shared memory'),
                                'regs': t.get('args', {}).get('registers per thread')}) + '\n')
    wall = one[-1]['ts'] + one[-1]['dur'] - t0
    streams = {}
    for t in one:
        streams.setdefault(t.get('args', {}).get('stream'), []).append(t)
    with open(args.out + '.txt', 'w') as f:
        f.write('step wall %.1f us, %d kernels/copies, %d streams\n' % (wall, len(one), len(streams)))
        for s, lst in sorted(streams.items(), key=lambda kv: -sum(t['dur'] for t in kv[1])):
            busy = sum(t['dur'] for t in lst)
            f.write('stream %s: %d launches, busy %.1f us (%.0f %% of the step), first %.1f last %.1f\n' %
                    (s, len(lst), busy, 100 * busy / wall, lst[0]['ts'] - t0, lst[-1]['ts'] + lst[-1]['dur'] - t0))
    os.remove(tmp)
    print(open(args.out + '.txt').read())


if __name__ == '__main__':
    main()
