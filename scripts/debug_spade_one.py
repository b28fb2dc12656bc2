"""Debug: isolate up_1.bn_1's SPADE backward on the tensors it sees inside the network."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
from fsv import networks, ops
from oracle import ops as O
from fsvtest import load_npz, state_from, opt_from, T, rel_err, grad_err
ops.CONV_USE_TC = 0
z = load_npz('g_face_tiny.npz')
opt = opt_from(z); opt.gpu_ids = [0]
G = networks.define_G(opt); G.load_state_dict(state_from(z, 'sd.')); G.train()
cap = {}
mod = G.up_1.bn_1
def pre(m, args, kwargs):
    cap['args'] = args; cap['kwargs'] = kwargs
def post(m, inp, out):
    out.retain_grad(); cap['out'] = out
mod.register_forward_pre_hook(pre, with_kwargs=True)
mod.register_forward_hook(post)
rm0, rv0 = mod.norm.running_mean.clone(), mod.norm.running_var.clone()
label, lref, iref = T(z['label']).cuda(), T(z['lref']).cuda(), T(z['iref']).cuda()
out = G(label, lref, iref)
(out[0] * T(z['r1']).cuda()).sum().backward()
x, maps, weights = cap['args'][0], cap['args'][1], cap['args'][2]
kw = cap['kwargs']
dout = cap['out'].grad.detach()
print('x', tuple(x.shape), 'maps', [None if m is None else tuple(m.shape) for m in maps], 'kw', {k: v for k, v in kw.items()})
# standalone rerun of my module
mod.norm.running_mean.copy_(rm0); mod.norm.running_var.copy_(rv0)
xg = x.detach().clone().requires_grad_(True)
mg = [None if m is None else m.detach().clone().requires_grad_(True) for m in maps]
flat, o1, o2, o3, o4 = weights
fg = flat.detach().clone().requires_grad_(True)
y = mod(xg, mg, (fg, o1, o2, o3, o4), **kw)
(y * dout).sum().backward()
# oracle fp64 on identical tensors
dt = torch.float64
V = lambda t: t.detach().permute(0, 3, 1, 2).cpu().to(dt)
xc = V(x).requires_grad_(True)
mc = [None if m is None else V(m).requires_grad_(True) for m in maps]
fc = flat.detach().cpu().to(dt).requires_grad_(True)
C, K = x.shape[3], maps[0].shape[3]
wts = O.slice_gamma_beta(fc, [C, K, 1, 1])
sd = {'s.' + k: v.detach().cpu().to(dt).clone() for k, v in mod.state_dict().items()}
sd['s.norm.running_mean'] = rm0.cpu().to(dt); sd['s.norm.running_var'] = rv0.cpu().to(dt)
yc = O.spade(xc, mc, sd, 's', 'batch', True, wts)
yc = O.lrelu(yc) if kw.get('act', 0) == 1 else yc
(yc * V(dout)).sum().backward()
a = xg.grad.permute(0, 3, 1, 2).cpu().to(dt); b = xc.grad
print('fwd err', rel_err(y.permute(0, 3, 1, 2), yc), 'dx err max', grad_err(a, b), 'l2', float((a - b).norm() / b.norm()))
d = (a - b).abs()
idx = torch.topk(d.flatten(), 8).indices
mean = xc.detach().mean((0, 2, 3)); var = xc.detach().var((0, 2, 3), unbiased=False)
xh = (xc.detach() - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5)
print('per-channel mean', mean.tolist()); print('per-channel std', var.sqrt().tolist())
for i in idx.tolist():
    n, c, h, w = (i // (C * 32 * 32)), (i // (32 * 32)) % C, (i // 32) % 32, i % 32
    print('n%d c%d h%d w%d: mine %.6e ref %.6e  xhat %.3f  pre-act %.3e' % (n, c, h, w, a.flatten()[i], b.flatten()[i], xh.flatten()[i], float(O.spade(xc.detach(), [m.detach() for m in mc if m is not None], sd, 's', 'batch', False, wts).flatten()[i]) if False else 0))
yv = O.spade(xc.detach(), [None if m is None else m.detach() for m in mc], dict(sd), 's', 'batch', True, [[w_.detach() for w_ in p_] for p_ in wts])
print('min |pre-activation| =', float(yv.abs().min()), ' count |v|<1e-5:', int((yv.abs() < 1e-5).sum()), 'of', yv.numel())
for i in idx.tolist():
    print('  pre-act at worst idx: %.3e' % float(yv.flatten()[i]))
