#!/usr/bin/env python
"""Micro-driver for the thin-layer kernels (profiling aid): one forward + backward of each thin shape of the face-256 step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'few-shot-vid2vid_b200'))
from fsv import ops  # noqa: E402

SHAPES = [  # N, H, W, Cin, Cout, k, stride, pad, in_act
    (8, 256, 256, 32, 3, 3, 1, 1, 1), (8, 256, 256, 32, 1, 3, 1, 1, 0), (16, 18, 18, 512, 1, 4, 1, 2, 0),
    (8, 256, 256, 1, 32, 3, 1, 1, 0), (8, 256, 256, 5, 32, 3, 1, 1, 0), (16, 256, 256, 8, 32, 4, 2, 2, 0),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for (n, h, w, ci, co, k, s, p, ia) in SHAPES:
    x = torch.randn(n, h, w, ci, device='cuda').requires_grad_(True)
    wt = (torch.randn(co, k, k, ci, device='cuda') * 0.1).requires_grad_(True)
    b = torch.zeros(co, device='cuda', requires_grad=True)
    for it in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = ops.conv2d(x, wt, b, stride=s, pad=p, in_act=ia)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        y.backward(torch.ones_like(y))
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print('%s fwd %.1f us, bwd %.1f us' % ((n, h, w, ci, co, k, s), (t1 - t0) * 1e6, (t2 - t1) * 1e6))
