"""Debug: does autograd deliver SpadeFn.backward's dx unchanged to the producer of x?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
from fsv import networks, ops
from util import load_npz, state_from, opt_from, T
ops.CONV_USE_TC = 0
z = load_npz('g_face_tiny.npz')
opt = opt_from(z); opt.gpu_ids = [0]
G = networks.define_G(opt); G.load_state_dict(state_from(z, 'sd.')); G.train()
rec = {}
orig_bwd = ops.SpadeFn.backward
def wrapped(ctx, dout):
    res = orig_bwd(ctx, dout)
    if ctx.dims == (2, 32, 32, 8):
        rec['dout'] = dout.detach().clone(); rec['dx'] = res[0].detach().clone(); rec['dx_obj'] = res[0]
        rec['dout_contig'] = dout.is_contiguous(); rec['dout_strides'] = dout.stride()
    return res
ops.SpadeFn.backward = staticmethod(wrapped)
cap = {}
def post(name):
    def f(m, inp, out):
        out.retain_grad(); cap[name] = out
    return f
G.up_1.conv_0.register_forward_hook(post('conv_0')); G.up_1.bn_1.register_forward_hook(post('bn_1'))
label, lref, iref = T(z['label']).cuda(), T(z['lref']).cuda(), T(z['iref']).cuda()
out = G(label, lref, iref)
(out[0] * T(z['r1']).cuda()).sum().backward()
print('dout contiguous', rec['dout_contig'], rec['dout_strides'])
print('dout seen by backward == retained grad of bn_1 out :', float((rec['dout'] - cap['bn_1'].grad).abs().max()))
print('dx returned == retained grad of conv_0 out        :', float((rec['dx'] - cap['conv_0'].grad).abs().max()), 'scale', float(rec['dx'].abs().max()))
print('dx object after backward == clone at return        :', float((rec['dx'] - rec['dx_obj']).abs().max()))
print('conv_0 out requires_grad', cap['conv_0'].requires_grad, 'is_leaf', cap['conv_0'].is_leaf, 'grad_fn', cap['conv_0'].grad_fn)
