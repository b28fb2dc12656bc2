"""Debug helper: stage-by-stage comparison of the CUDA generator against the CPU oracle (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'few-shot-vid2vid_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
from fsv import networks, ops
from oracle import nets as ON
from util import load_npz, state_from, opt_from, T, rel_err

ops.CONV_USE_TC = 0
z = load_npz('g_face_tiny.npz')
opt = opt_from(z); opt.gpu_ids = [0]
G = networks.define_G(opt)
G.load_state_dict(state_from(z, 'sd.'))
G.train()
cap = {}
def hook(name):
    def f(mod, inp, out):
        cap[name] = out
    return f
for i in range(opt.n_downsample_G + 1):
    getattr(G, 'up_%d' % i).register_forward_hook(hook('up_%d' % i))
G.label_embedding.register_forward_hook(hook('label_embedding'))
G.img_ref_embedding.register_forward_hook(hook('img_ref_embedding'))
for i in range(opt.n_downsample_G):
    getattr(G, 'ref_img_down_%d' % i).register_forward_hook(hook('ref_img_down_%d' % i))
G.ref_img_first.register_forward_hook(hook('ref_img_first'))
label, lref, iref = T(z['label']).cuda(), T(z['lref']).cuda(), T(z['iref']).cuda()
out = G(label, lref, iref)
sd = state_from(z, 'sd.')
ref, internals = ON.generator_forward(sd, opt_from(z), T(z['label']), T(z['lref']), T(z['iref']), training=True, return_internals=True)
V = lambda t: t.permute(0, 3, 1, 2)
print('flow', rel_err(out[1][0], ref[1][0]), 'mask', rel_err(out[2][0], ref[2][0]), 'warp', rel_err(out[4][0], ref[4][0]), 'img', rel_err(out[0], ref[0]))
for i, (a, b) in enumerate(zip(cap['label_embedding'], [e[0] if isinstance(e, list) else e for e in internals['enc_label']])):
    print('label_emb', i, tuple(a.shape), rel_err(V(a), b))
for i, a in enumerate(cap['img_ref_embedding']):
    e = internals['enc_label'][i]
    if isinstance(e, list):
        print('img_ref_emb', i, rel_err(V(a), e[1]))
for i in range(opt.n_downsample_G, -1, -1):
    print('up_%d' % i, rel_err(V(cap['up_%d' % i]), internals['up_%d' % i]))
# hyper weights
for i, nw in enumerate(internals['norm_w']):
    mine = G.get_SPADE_weights  # noqa
