"""Debug helper: per-parameter gradient comparison CUDA generator vs CPU oracle (fp64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
from fsv import networks, ops
from oracle import nets as ON
from util import load_npz, state_from, opt_from, T, rel_err, grad_err

ops.CONV_USE_TC = 0
z = load_npz('g_face_tiny.npz')
opt = opt_from(z); opt.gpu_ids = [0]
G = networks.define_G(opt)
G.load_state_dict(state_from(z, 'sd.'))
G.train()
label, lref, iref = T(z['label']).cuda(), T(z['lref']).cuda(), T(z['iref']).cuda()
out = G(label, lref, iref)
C = lambda a: T(a).cuda()
loss = ((out[0] * C(z['r1'])).sum() + 0.05 * (out[1][0] * C(z['r2'])).sum() + (out[2][0] * C(z['r3'])).sum() + (out[4][0] * C(z['r4'])).sum())
loss.backward()
dt = torch.float64
sd = state_from(z, 'sd.', dtype=dt)
for k, v in sd.items():
    if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
        v.requires_grad_(True)
ref = ON.generator_forward(sd, opt_from(z), T(z['label'], dt), T(z['lref'], dt), T(z['iref'], dt), training=True)
l2 = ((ref[0] * T(z['r1'], dt)).sum() + 0.05 * (ref[1][0] * T(z['r2'], dt)).sum() + (ref[2][0] * T(z['r3'], dt)).sum() + (ref[4][0] * T(z['r4'], dt)).sum())
l2.backward()
rows = []
for n, p in G.named_parameters():
    if p.grad is None or sd[n].grad is None:
        rows.append((9.9, n, 'missing grad mine=%s ref=%s' % (p.grad is not None, sd[n].grad is not None)))
        continue
    a, b = p.grad.detach().cpu().double(), sd[n].grad
    e = grad_err(a, b)
    rl2 = float((a - b).norm() / (b.norm() + 1e-30))
    rows.append((e, n, 'max %.2e l2 %.2e |ref|max %.2e' % (e, rl2, float(b.abs().max()))))
rows.sort(reverse=True)
for e, n, s in rows[:60]:
    print('%-55s %s' % (n, s))
print('... %d params total, %d with err > 1e-3' % (len(rows), sum(1 for r in rows if r[0] > 1e-3)))

# ---- second pass: activation gradients at block outputs, and per-module summary
print('=== per top-level module: max grad err')
agg = {}
for e, n, s in rows:
    if e < 9:
        top = n.split('.')[0]
        agg[top] = max(agg.get(top, 0.0), e)
for k in sorted(agg, key=lambda k: -agg[k]):
    print('%-30s %.2e' % (k, agg[k]))

G.zero_grad()
G.load_state_dict(state_from(z, 'sd.'))   # training forward advanced u/v and running stats: reset
cap = {}
def hook(name):
    def f(mod, inp, out):
        out.retain_grad(); cap[name] = out
    return f
names = []
for i in range(opt.n_downsample_G + 1):
    blk = getattr(G, 'up_%d' % i)
    blk.register_forward_hook(hook('up_%d' % i)); names.append('up_%d' % i)
    for sub in ('bn_0', 'conv_0', 'bn_1', 'conv_s'):
        if hasattr(blk, sub):
            getattr(blk, sub).register_forward_hook(hook('up_%d.%s' % (i, sub))); names.append('up_%d.%s' % (i, sub))
out = G(label, lref, iref)
loss = (out[0] * C(z['r1'])).sum()
loss.backward()
sd = state_from(z, 'sd.', dtype=dt)
for k, v in sd.items():
    if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
        v.requires_grad_(True)
ref, internals = ON.generator_forward(sd, opt_from(z), T(z['label'], dt), T(z['lref'], dt), T(z['iref'], dt), training=True, return_internals=True)
for n in names:
    if n in internals and internals[n].requires_grad:
        internals[n].retain_grad()
(ref[0] * T(z['r1'], dt)).sum().backward()
for n in names:
    if n not in internals or internals[n].grad is None or cap[n].grad is None:
        print('skip', n); continue
    a = cap[n].grad.permute(0, 3, 1, 2).cpu().double(); b = internals[n].grad
    print('d(%-14s): max %.2e l2 %.2e | fwd err %.2e | sum-per-channel mine %.2e ref %.2e' % (n, grad_err(a, b), float((a - b).norm() / b.norm()),
          rel_err(cap[n].permute(0, 3, 1, 2), internals[n]), float(a.sum((0, 2, 3)).abs().max()), float(b.sum((0, 2, 3)).abs().max())))
