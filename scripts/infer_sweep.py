#!/usr/bin/env python
"""BASELINE.json config 5: generator-only inference sweep (eval mode, no_grad, face geometry) over frame size and number of
reference images K, for the fsv drop-in generator and the unmodified reference generator (baseline/_ref, cuDNN TF32) on the same GPU.

Per (size, K): frame 0 (reference encoding + hyper-network weight generation + synthesis) and the steady state t >= 1 (cached
weights, previous-frame warp branch: vid2vid_model.py:179-205, generator.py:403-418), frames/s = 1 / time per frame at batch 1.
Writes one JSON line per row to --out; rows that run out of memory or exceed the time limit are recorded as such."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200'), os.path.join(ROOT, 'baseline')):
    sys.path.insert(0, p)


def child(impl, size, K, frames):
    import torch
    import refenv
    import synth
    opt = refenv.parse_opt('face', size, size, 1, extra=['--n_shot', str(K)] if K > 1 else [], gpu=True, train=False)
    opt.isTrain = False
    opt.for_face = False                      # set by BaseModel.define_networks in the reference (base_model.py:172)
    if impl == 'fsv':
        from fsv import networks
    else:
        import models.networks as networks
        torch.backends.cudnn.benchmark = True
    G = networks.define_G(opt)
    G.init_temporal_network()
    G.cuda()
    b = {k: v.cuda() for k, v in synth.make('face', 1, size, size, seed=3, K=K).items()}
    label, lref, iref = b['tgt_label'][:, 0], b['ref_label'], b['ref_image']
    G.train()       # BatchNorm running statistics of a fresh network are (0, 1): populate them (both arms) so that eval-mode values are finite
    with torch.no_grad():
        for _ in range(6):
            G(label, lref, iref, [label, b['tgt_image'][:, 0]])
    G.eval()

    def run(n):
        prev = [None, None]
        ts = []
        for t in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.no_grad():
                out = G(label, lref, iref, prev, t=t)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            prev = [label, out[0].contiguous()]
        return ts
    torch.cuda.set_stream(torch.cuda.Stream())
    run(3)                  # warm-up (cudnn.benchmark, allocator, plans)
    ts = run(frames)
    steady = sorted(ts[1:])[len(ts[1:]) // 2]
    rec = dict(impl=impl, size=size, K=K, first_frame_ms=ts[0] * 1e3, steady_ms=steady * 1e3, steady_fps=1.0 / steady)
    if impl == 'fsv':       # the same steady-state frame replayed from a CUDA graph (fsv.infer.GraphedGenerator)
        from fsv.infer import GraphedGenerator
        with torch.no_grad():
            out = G(label, lref, iref, [None, None], t=0)
            prev = [label, out[0].contiguous()]
        gg = GraphedGenerator(G, label, lref, iref, prev)
        tg = []
        for _ in range(frames):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o = gg(label, prev)
            torch.cuda.synchronize()
            tg.append(time.perf_counter() - t0)
            prev = [label, o[0].contiguous()]
        g = sorted(tg)[len(tg) // 2]
        rec.update(graph_steady_ms=g * 1e3, graph_steady_fps=1.0 / g)
    rec['peak_mem_gb'] = torch.cuda.max_memory_allocated() / 2 ** 30
    print('RESULT ' + json.dumps(rec), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'infer_sweep.jsonl'))
    ap.add_argument('--sizes', default='256,512,1024')
    ap.add_argument('--shots', default='1,5,20')
    ap.add_argument('--frames', type=int, default=12)
    ap.add_argument('--child', nargs=3, default=None)
    a = ap.parse_args()
    if a.child:
        child(a.child[0], int(a.child[1]), int(a.child[2]), a.frames)
        sys.exit(0)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    for size in [int(x) for x in a.sizes.split(',')]:
        for K in [int(x) for x in a.shots.split(',')]:
            for impl in ('fsv', 'reference'):
                try:
                    p = subprocess.run([sys.executable, __file__, '--child', impl, str(size), str(K), '--frames', str(a.frames)],
                                       capture_output=True, text=True, timeout=420)
                    line = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')]
                    rec = json.loads(line[-1][7:]) if line else dict(impl=impl, size=size, K=K, error=('out of memory' if 'out of memory' in p.stderr else p.stderr[-300:]))
                except subprocess.TimeoutExpired:
                    rec = dict(impl=impl, size=size, K=K, error='timeout 420 s')
                with open(a.out, 'a') as f:
                    f.write(json.dumps(rec) + '\n')
                print(json.dumps(rec), flush=True)
