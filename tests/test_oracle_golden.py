"""Pin the CPU oracle against the reference's own outputs (golden fixtures made by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ops, nets
from fsvtest import load_npz, state_from, opt_from, T, rel_err, grad_err

TOL = 2e-5   # fp32 oracle vs fp32 reference: same math, different op order


def test_reshape_weight_layout():
    z = load_npz('ops.npz')
    gb = ops.slice_gamma_beta(T(z['reshape.flat']), [5, 3, 1, 1])
    assert torch.equal(gb[0][0], T(z['reshape.gw'])) and torch.equal(gb[0][1], T(z['reshape.gb']))
    assert torch.equal(gb[1][0], T(z['reshape.bw'])) and torch.equal(gb[1][1], T(z['reshape.bb']))
    fe = T(z['reshape.flat_e'])
    ew = ops.slice_weight_bias(fe[:, :-3], [3, 6, 1, 1])
    assert torch.equal(ew[0], T(z['reshape.ew'])) and torch.equal(ew[1], T(z['reshape.eb']))


@pytest.mark.parametrize('kind', ['batch', 'instance'])
def test_spade_op(kind):
    z = load_npz('ops.npz')
    pre = 'spade_%s.' % kind
    sd = {'s.' + k: v for k, v in state_from(z, pre + 'sd.').items()}
    maps = [T(z[pre + 'm0']), T(z[pre + 'm1']), T(z[pre + 'm2'])]
    wts = [[[T(z[pre + 'wg']), T(z[pre + 'bg'])]], [[T(z[pre + 'wb']), T(z[pre + 'bb'])]]]
    y = ops.spade(T(z[pre + 'x']), maps, sd, 's', kind, True, wts)
    assert rel_err(y, T(z[pre + 'y'])) < TOL


def test_spade_eval_fixed():
    z = load_npz('ops.npz')
    sd = {'s.' + k: v for k, v in state_from(z, 'spade_eval.sd.').items()}
    y = ops.spade(T(z['spade_eval.x']), T(z['spade_eval.m0']), sd, 's', 'batch', False, None)
    assert rel_err(y, T(z['spade_eval.y'])) < TOL


def test_warp_bconv_outer():
    z = load_npz('ops.npz')
    assert rel_err(ops.resample(T(z['warp.img']), T(z['warp.flow'])), T(z['warp.out'])) < TOL
    assert rel_err(ops.batch_conv(T(z['bconv.x']), T(z['bconv.w']), T(z['bconv.b'])), T(z['bconv.y'])) < TOL
    assert rel_err(ops.ref_outer_product(T(z['outer.a']), T(z['outer.l'])), T(z['outer.y'])) < TOL


def _g_loss(out, z):
    return ((out[0] * T(z['r1'])).sum() + 0.05 * (out[1][0] * T(z['r2'])).sum() +
            (out[2][0] * T(z['r3'])).sum() + (out[4][0] * T(z['r4'])).sum())


def test_generator_train_forward_backward():
    z = load_npz('g_face_tiny.npz')
    opt = opt_from(z)
    sd = state_from(z, 'sd.')
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
            v.requires_grad_(True)
    label = T(z['label']).requires_grad_(True)
    iref = T(z['iref']).requires_grad_(True)
    out = nets.generator_forward(sd, opt, label, T(z['lref']), iref, training=True)
    assert out[3] is None and out[1][1] is None
    assert rel_err(out[0], T(z['out_img'])) < TOL
    assert rel_err(out[1][0], T(z['out_flow'])) < TOL
    assert rel_err(out[2][0], T(z['out_mask'])) < TOL
    assert rel_err(out[4][0], T(z['out_warp'])) < TOL
    loss = _g_loss(out, z)
    assert abs(loss.item() - float(z['loss'])) < 1e-3 * abs(float(z['loss'])) + 1e-3
    loss.backward()
    for k in z.files:
        if k.startswith('grad.'):
            assert grad_err(sd[k[5:]].grad, T(z[k])) < 2e-4, k
    assert rel_err(label.grad, T(z['grad_label'])) < 2e-4
    assert rel_err(iref.grad, T(z['grad_iref'])) < 2e-4
    for k in z.files:
        if k.startswith('post.'):
            assert rel_err(sd[k[5:]].detach().float(), T(z[k])) < TOL, k


def test_generator_eval_with_weight_cache():
    z = load_npz('g_face_tiny_eval.npz')
    opt = opt_from(z)
    sd = state_from(z, 'sd.')
    with torch.no_grad():
        o0, internals = nets.generator_forward(sd, opt, T(z['label0']), T(z['lref']), T(z['iref']),
                                               training=False, return_internals=True)
        assert rel_err(o0[0], T(z['out_img0'])) < TOL
        cache = (internals['emb_w'], internals['norm_w'])
        o1 = nets.generator_forward(sd, opt, T(z['label1']), T(z['lref']), T(z['iref']), training=False,
                                    cached_weights=cache)
    assert rel_err(o1[0], T(z['out_img1'])) < TOL
    assert rel_err(o1[1][0], T(z['out_flow1'])) < TOL
    assert rel_err(o1[2][0], T(z['out_mask1'])) < TOL


def test_generator_temporal():
    z = load_npz('g_face_tiny_temporal.npz')
    opt = opt_from(z)
    sd = state_from(z, 'sd.')
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
            v.requires_grad_(True)
    out = nets.generator_forward(sd, opt, T(z['label']), T(z['lref']), T(z['iref']),
                                 prev=(T(z['prev_label']), T(z['prev_img'])), training=True, temporal=True)
    assert rel_err(out[0], T(z['out_img'])) < TOL
    assert rel_err(out[1][1], T(z['out_flow_prev'])) < TOL
    assert rel_err(out[2][1], T(z['out_mask_prev'])) < TOL
    assert rel_err(out[4][1], T(z['out_warp_prev'])) < TOL
    loss = (out[0] * T(z['r1'])).sum() + (out[4][1] * T(z['r4'])).sum() + (out[2][1] * T(z['r3'])).sum()
    loss.backward()
    for k in z.files:
        if k.startswith('grad.'):
            g = sd[k[5:]].grad   # flow_network_temp IS flow_network_ref here: its grads add up in one tensor
            assert grad_err(g, T(z[k])) < 2e-4, k


def test_discriminator():
    z = load_npz('d_tiny.npz')
    sd = state_from(z, 'sd.')
    for k, v in sd.items():
        if not k.endswith(('weight_u', 'weight_v')):
            v.requires_grad_(True)
    x = T(z['x']).requires_grad_(True)
    pred = nets.discriminator_forward(sd, x, n_layers=4, num_D=2, training=True)
    loss = 0
    for i, p in enumerate(pred):
        assert len(p) == 6
        for j, t in enumerate(p):
            assert rel_err(t, T(z['out.%d.%d' % (i, j)])) < TOL, (i, j)
            loss = loss + (t * T(z['r.%d.%d' % (i, j)])).sum()
    loss.backward()
    assert rel_err(x.grad, T(z['grad_x'])) < 2e-4
    for k in z.files:
        if k.startswith('grad.'):
            assert grad_err(sd[k[5:]].grad, T(z[k])) < 2e-4, k
        if k.startswith('post.'):
            assert rel_err(sd[k[5:]].detach(), T(z[k])) < TOL, k


def test_train_step_losses():
    z = load_npz('step_face_tiny.npz')
    zg = load_npz('g_face_tiny.npz')
    opt = opt_from(zg)
    sdG = state_from(zg, 'sd.')
    sdD = state_from(z, 'sdD.')
    for sd in (sdG, sdD):
        for k, v in sd.items():
            if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
                v.requires_grad_(True)
    label, lref, iref, tgt = T(z['label']), T(z['lref']), T(z['iref']), T(z['tgt'])
    with torch.no_grad():
        fake_d = nets.generator_forward(sdG, opt, label, lref, iref, training=True)[0]
    dl = nets.discriminator_losses(sdD, label, fake_d, tgt, lref[:, 0], iref[:, 0])
    assert rel_err(dl['D_real'], T(z['D_real'])) < 1e-4 and rel_err(dl['D_fake'], T(z['D_fake'])) < 1e-4
    sum(dl.values()).sum().backward()
    for k in z.files:
        if k.startswith('gradD.'):
            assert grad_err(sdD[k[6:]].grad, T(z[k])) < 3e-4, k
    gl, fake = nets.generator_losses(sdG, sdD, opt, label, tgt, lref, iref)
    assert rel_err(fake, T(z['fake'])) < 1e-4
    for n in ('G_GAN', 'G_GAN_Feat', 'F_Warp', 'F_Mask'):
        assert rel_err(gl[n].reshape(-1), T(z[n]).reshape(-1)) < 1e-4, n
    assert float(np.abs(z['F_Flow']).max()) == 0.0
    sum(v.sum() for v in gl.values()).backward()
    for k in z.files:
        if k.startswith('gradG.'):
            assert grad_err(sdG[k[6:]].grad, T(z[k])) < 3e-4, k


@pytest.mark.parametrize('name', ['pose', 'street'])
def test_generator_other_dataset_geometries(name):
    """Oracle vs the reference on two more BASELINE geometries (tests/golden/make_golden.py variants): 'pose' = 6-channel
    label, portrait H = 2W, warp_ref + spade_combine; 'street' = wide W = 2H, --adaptive_spade only (no flow branch)."""
    import json
    from argparse import Namespace
    z = load_npz('g_variants_tiny.npz')
    pre = name + '.'
    opt = Namespace(**json.loads(str(z[pre + 'opt'])))
    sd = state_from(z, pre + 'sd.')
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
            v.requires_grad_(True)
    label = T(z[pre + 'label']).requires_grad_(True)
    out = nets.generator_forward(sd, opt, label, T(z[pre + 'lref']), T(z[pre + 'iref']), training=True)
    assert rel_err(out[0], T(z[pre + 'out_img'])) < TOL
    loss = (out[0] * T(z[pre + 'r1'])).sum()
    if int(z[pre + 'has_flow']):
        assert rel_err(out[1][0], T(z[pre + 'out_flow'])) < TOL
        assert rel_err(out[2][0], T(z[pre + 'out_mask'])) < TOL
        loss = loss + 0.05 * out[1][0].sum() + out[2][0].sum()
    else:
        assert out[1][0] is None and out[2][0] is None
    assert abs(float(loss.detach()) - float(z[pre + 'loss'])) < 1e-3 * max(1.0, abs(float(z[pre + 'loss'])))
    loss.backward()
    assert grad_err(label.grad, T(z[pre + 'grad_label'])) < 1e-3
    for k in z.files:
        if k.startswith(pre + 'grad.') and k != pre + 'grad_label':
            n = k[len(pre) + 5:]
            assert grad_err(sd[n].grad, T(z[k])) < 1e-3, n


def test_generator_two_reference_images_attention():
    """K = 2: the oracle's attention module / ref_idx / pick_ref against the reference
    (tests/golden/g_kshot_tiny.npz): frames, flow, mask, warp, the attention visualisation, the picked reference and
    gradients -- including those of the attention key / query encoders."""
    z = load_npz('g_kshot_tiny.npz')
    opt = opt_from(z)
    sd = state_from(z, 'sd.')
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
            v.requires_grad_(True)
    label = T(z['label']).requires_grad_(True)
    out = nets.generator_forward(sd, opt, label, T(z['lref']), T(z['iref']), training=True)
    assert rel_err(out[0], T(z['out_img'])) < TOL
    assert rel_err(out[1][0], T(z['out_flow'])) < TOL
    assert rel_err(out[2][0], T(z['out_mask'])) < TOL
    assert rel_err(out[4][0], T(z['out_warp'])) < TOL
    assert rel_err(out[7], T(z['atn_vis'])) < TOL
    assert torch.equal(out[8], torch.from_numpy(np.array(z['ref_idx'])))
    loss = (out[0] * T(z['r1'])).sum() + 0.05 * out[1][0].sum() + out[2][0].sum()
    loss.backward()
    assert grad_err(label.grad, T(z['grad_label'])) < 1e-3
    for k in z.files:
        if k.startswith('grad.'):
            assert grad_err(sd[k[5:]].grad, T(z[k])) < 1e-3, k


def test_generator_losses_temporal_phase():
    """oracle generator_losses with a previous frame against the reference LossCollector (step_face_tiny_temporal.npz)."""
    z = load_npz('step_face_tiny_temporal.npz')
    zt = load_npz('g_face_tiny_temporal.npz')
    opt = opt_from(zt)
    losses, fake = nets.generator_losses(state_from(zt, 'sd.'), state_from(z, 'sdD.'), opt, T(zt['label']), T(z['tgt']),
                                         T(zt['lref']), T(zt['iref']), prev=(T(zt['prev_label']), T(zt['prev_img'])))
    assert rel_err(fake, T(z['fake'])) < TOL
    for n in ('G_GAN', 'G_GAN_Feat', 'F_Warp', 'F_Mask'):
        assert rel_err(losses[n].reshape(-1), T(z[n]).reshape(-1)) < TOL, n
