"""Host-logic tests on CPU: the drop-in modules of ``fsv.networks`` (the very code that runs on the GPU) with ``fsv.ops``
replaced by the torch-CPU emulation in tests/mock_ops.py, against the golden fixtures produced by the REFERENCE.  This checks
everything the Python host side decides -- which op gets which tensors, hyper-weight offsets inside the flat MLP outputs,
the reference's quirks (no hyper bias in adaptive SPADE, shared flow net called twice), state handling (BN / spectral buffers,
eval weight cache, temporal phase) and the K-shot attention wiring -- without a GPU.  The kernels themselves are checked by the
``-m gpu`` tests; nothing in the product imports the emulation."""
import json
from argparse import Namespace

import numpy as np
import pytest
import torch

import mock_ops
from fsvtest import load_npz, state_from, opt_from, T, rel_err, l2_err

TOL = 5e-5
GTOL = 1e-2     # network-level gradients: relative L2 (a LeakyReLU kink flip perturbs single entries, see util.l2_err)


def grad_err(a, b):
    return l2_err(a, b)


@pytest.fixture()
def nets(monkeypatch):
    from fsv import networks
    from fsv.networks import layers, generator, discriminator
    for mod in (layers, generator, discriminator):
        monkeypatch.setattr(mod, 'ops', mock_ops)
    return networks


def _build(nets, opt, sd, temporal=False, train=True):
    opt.gpu_ids = []
    G = nets.define_G(opt)
    if temporal:
        G.init_temporal_network()
    G.load_state_dict(sd)
    G.train(train)
    return G


def test_generator_train_forward_backward(nets):
    z = load_npz('g_face_tiny.npz')
    G = _build(nets, opt_from(z), state_from(z, 'sd.'))
    label = T(z['label']).requires_grad_(True)
    iref = T(z['iref']).requires_grad_(True)
    out = G(label, T(z['lref']), iref)
    assert out[3] is None and out[1][1] is None and out[7] is None and out[8] is None
    assert rel_err(out[0], T(z['out_img'])) < TOL
    assert rel_err(out[1][0], T(z['out_flow'])) < TOL
    assert rel_err(out[2][0], T(z['out_mask'])) < TOL
    assert rel_err(out[4][0], T(z['out_warp'])) < TOL
    loss = ((out[0] * T(z['r1'])).sum() + 0.05 * (out[1][0] * T(z['r2'])).sum() + (out[2][0] * T(z['r3'])).sum() +
            (out[4][0] * T(z['r4'])).sum())
    loss.backward()
    params = dict(G.named_parameters())
    for k in z.files:
        if k.startswith('grad.'):
            assert grad_err(params[k[5:]].grad, T(z[k])) < GTOL, k
    assert grad_err(label.grad, T(z['grad_label'])) < GTOL
    assert grad_err(iref.grad, T(z['grad_iref'])) < GTOL
    sd1 = G.state_dict()
    for k in z.files:
        if k.startswith('post.'):          # BN running stats, spectral u / v, counters after one training forward
            assert rel_err(sd1[k[5:]].float(), T(z[k])) < TOL, k


def test_generator_eval_weight_cache(nets):
    z = load_npz('g_face_tiny_eval.npz')
    G = _build(nets, opt_from(z), state_from(z, 'sd.'), train=False)
    with torch.no_grad():
        o0 = G(T(z['label0']), T(z['lref']), T(z['iref']), t=0)
        o1 = G(T(z['label1']), T(z['lref']), T(z['iref']), t=1)
    assert rel_err(o0[0], T(z['out_img0'])) < TOL
    assert rel_err(o1[0], T(z['out_img1'])) < TOL
    assert rel_err(o1[1][0], T(z['out_flow1'])) < TOL
    assert rel_err(o1[2][0], T(z['out_mask1'])) < TOL


def test_generator_temporal_phase(nets):
    z = load_npz('g_face_tiny_temporal.npz')
    G = _build(nets, opt_from(z), state_from(z, 'sd.'), temporal=True)
    out = G(T(z['label']), T(z['lref']), T(z['iref']), prev=[T(z['prev_label']), T(z['prev_img'])])
    assert rel_err(out[0], T(z['out_img'])) < TOL
    assert rel_err(out[1][1], T(z['out_flow_prev'])) < TOL
    assert rel_err(out[2][1], T(z['out_mask_prev'])) < TOL
    assert rel_err(out[4][1], T(z['out_warp_prev'])) < TOL
    loss = (out[0] * T(z['r1'])).sum() + (out[4][1] * T(z['r4'])).sum() + (out[2][1] * T(z['r3'])).sum()
    assert abs(float(loss.detach()) - float(z['loss'])) < 1e-3 * max(1.0, abs(float(z['loss'])))
    loss.backward()
    params = dict(G.named_parameters())
    for k in z.files:
        if k.startswith('grad.'):
            assert grad_err(params[k[5:]].grad, T(z[k])) < GTOL, k


def test_discriminator(nets):
    z = load_npz('d_tiny.npz')
    zg = load_npz('g_face_tiny.npz')
    opt = opt_from(zg)
    D = nets.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 2, True, gpu_ids=[])
    D.load_state_dict(state_from(z, 'sd.'))
    D.train()
    x = T(z['x']).requires_grad_(True)
    pred = D(x)
    loss = 0
    for i, p in enumerate(pred):
        for j, t in enumerate(p):
            assert rel_err(t, T(z['out.%d.%d' % (i, j)])) < TOL, (i, j)
            loss = loss + (t * T(z['r.%d.%d' % (i, j)])).sum()
    loss.backward()
    assert grad_err(x.grad, T(z['grad_x'])) < GTOL


@pytest.mark.parametrize('name', ['pose', 'street'])
def test_generator_other_dataset_geometries(nets, name):
    z = load_npz('g_variants_tiny.npz')
    pre = name + '.'
    opt = Namespace(**json.loads(str(z[pre + 'opt'])))
    G = _build(nets, opt, state_from(z, pre + 'sd.'))
    label = T(z[pre + 'label']).requires_grad_(True)
    out = G(label, T(z[pre + 'lref']), T(z[pre + 'iref']))
    assert rel_err(out[0], T(z[pre + 'out_img'])) < TOL
    loss = (out[0] * T(z[pre + 'r1'])).sum()
    if int(z[pre + 'has_flow']):
        assert rel_err(out[1][0], T(z[pre + 'out_flow'])) < TOL
        assert rel_err(out[2][0], T(z[pre + 'out_mask'])) < TOL
        loss = loss + 0.05 * out[1][0].sum() + out[2][0].sum()
    else:
        assert out[1][0] is None and out[2][0] is None
    loss.backward()
    assert grad_err(label.grad, T(z[pre + 'grad_label'])) < GTOL
    params = dict(G.named_parameters())
    for k in z.files:
        if k.startswith(pre + 'grad.') and k != pre + 'grad_label':
            assert grad_err(params[k[len(pre) + 5:]].grad, T(z[k])) < GTOL, k


def test_generator_two_reference_images(nets):
    z = load_npz('g_kshot_tiny.npz')
    G = _build(nets, opt_from(z), state_from(z, 'sd.'))
    label = T(z['label']).requires_grad_(True)
    out = G(label, T(z['lref']), T(z['iref']))
    assert rel_err(out[0], T(z['out_img'])) < TOL
    assert rel_err(out[1][0], T(z['out_flow'])) < TOL
    assert rel_err(out[2][0], T(z['out_mask'])) < TOL
    assert rel_err(out[4][0], T(z['out_warp'])) < TOL
    assert rel_err(out[7], T(z['atn_vis'])) < TOL
    assert torch.equal(out[8], torch.from_numpy(np.array(z['ref_idx'])))
    loss = (out[0] * T(z['r1'])).sum() + 0.05 * out[1][0].sum() + out[2][0].sum()
    loss.backward()
    assert grad_err(label.grad, T(z['grad_label'])) < GTOL
    params = dict(G.named_parameters())
    for k in z.files:
        if k.startswith('grad.'):
            assert grad_err(params[k[5:]].grad, T(z[k])) < GTOL, k


def test_generator_two_reference_images_chunked_attention(nets):
    """No-grad K = 2 forward with the attention matrix formed a few query rows at a time (FewShotGenerator.attention_chunked: the
    memory-bounded form the inference sweep needs at 1024x1024 / K = 5) against the one-piece form: frame, flow, warp, the attention
    visualisation and the picked reference."""
    z = load_npz('g_kshot_tiny.npz')
    sd = state_from(z, 'sd.')
    G = _build(nets, opt_from(z), sd)
    label, lref, iref = T(z['label']), T(z['lref']), T(z['iref'])
    outs = []
    with torch.no_grad():
        for budget in (1 << 40, 1):          # one piece; one query row per chunk
            G.load_state_dict(sd)                # same statistics / power-iteration state for both runs
            G.attention_chunk_bytes = budget
            nd = G.n_downsample_A
            rows = G.attention_rows_per_chunk(label.shape[0], label.shape[2] >> nd, label.shape[3] >> nd, G.n_shot)
            assert (rows is None) == (budget > 1) and (budget > 1 or rows == 1)
            outs.append(G(label, lref, iref))
    a, b = outs
    assert rel_err(b[0], a[0]) < 1e-5 and rel_err(b[1][0], a[1][0]) < 1e-5 and rel_err(b[4][0], a[4][0]) < 1e-5
    assert rel_err(b[7], a[7]) < 1e-5 and torch.equal(a[8], b[8])


def test_train_step_losses(nets, monkeypatch):
    """one D-step + G-step through fsv.trainer (the mirror of vid2vid_model.py:62-128 + loss_collector.py) on the
    emulated op layer: loss values, the generated frame and parameter gradients against the reference's LossCollector."""
    from fsv import trainer, model
    monkeypatch.setattr(model, 'ops', mock_ops)
    z = load_npz('step_face_tiny.npz')
    zg = load_npz('g_face_tiny.npz')
    opt = opt_from(zg)
    G = _build(nets, opt, state_from(zg, 'sd.'))
    D = nets.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 1, True, gpu_ids=[])
    D.load_state_dict(state_from(z, 'sdD.'))
    D.train()
    label, lref, iref, tgt = T(z['label']), T(z['lref']), T(z['iref']), T(z['tgt'])
    dl = trainer.discriminator_losses(opt, G, D, label, tgt, lref, iref)
    assert rel_err(dl['D_real'].reshape(-1), T(z['D_real']).reshape(-1)) < TOL
    assert rel_err(dl['D_fake'].reshape(-1), T(z['D_fake']).reshape(-1)) < TOL
    sum(v.mean() for v in dl.values()).backward()
    pd = dict(D.named_parameters())
    for k in z.files:
        if k.startswith('gradD.'):
            assert grad_err(pd[k[6:]].grad, T(z[k])) < GTOL, k
    D.zero_grad()
    gl, fake = trainer.generator_losses(opt, G, D, label, tgt, lref, iref)
    assert rel_err(fake, T(z['fake'])) < TOL
    for n in ('G_GAN', 'G_GAN_Feat', 'F_Warp', 'F_Mask'):
        assert rel_err(gl[n].reshape(-1), T(z[n]).reshape(-1)) < TOL, n
    sum(v.mean() for v in gl.values()).backward()
    pg = dict(G.named_parameters())
    for k in z.files:
        if k.startswith('gradG.'):
            assert grad_err(pg[k[6:]].grad, T(z[k])) < GTOL, k


def test_train_step_losses_temporal_phase(nets, monkeypatch):
    """G-step losses with a previous frame (warp_prev): warp / mask terms of both branches, against the reference's
    LossCollector (tests/golden/step_face_tiny_temporal.npz; state and inputs of g_face_tiny_temporal.npz)."""
    from fsv import trainer, model
    monkeypatch.setattr(model, 'ops', mock_ops)
    z = load_npz('step_face_tiny_temporal.npz')
    zt = load_npz('g_face_tiny_temporal.npz')
    opt = opt_from(zt)
    G = _build(nets, opt, state_from(zt, 'sd.'), temporal=True)
    D = nets.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 1, True, gpu_ids=[])
    D.load_state_dict(state_from(z, 'sdD.'))
    D.train()
    gl, fake = trainer.generator_losses(opt, G, D, T(zt['label']), T(z['tgt']), T(zt['lref']), T(zt['iref']),
                                        prev=[T(zt['prev_label']), T(zt['prev_img'])])
    assert rel_err(fake, T(z['fake'])) < TOL
    for n in ('G_GAN', 'G_GAN_Feat', 'F_Warp', 'F_Mask'):
        assert rel_err(gl[n].reshape(-1), T(z[n]).reshape(-1)) < TOL, n
    sum(v.mean() for v in gl.values()).backward()
    pg = dict(G.named_parameters())
    for k in z.files:
        if k.startswith('gradG.'):
            assert grad_err(pg[k[6:]].grad, T(z[k])) < GTOL, k


@pytest.mark.parametrize('temporal', [False, True])
def test_train_step_runs_and_updates_parameters(nets, monkeypatch, temporal):
    """trainer.train_step (what bench.py calls every iteration: D-step + G-step incl. both Adam updates) end to end on the
    emulated op layer, single-frame and temporal phase: finite losses, parameters of G and D actually move."""
    from fsv import trainer, model
    monkeypatch.setattr(model, 'ops', mock_ops)
    zt = load_npz('g_face_tiny_temporal.npz' if temporal else 'g_face_tiny.npz')
    opt = opt_from(zt)
    G = _build(nets, opt, state_from(zt, 'sd.'), temporal=temporal)
    D = nets.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 1, True, gpu_ids=[])
    D.train()
    optG, optD = trainer.make_optimizers(opt, G, D)
    w0 = G.conv_img.weight.detach().clone()
    d0 = next(D.parameters()).detach().clone()
    g = torch.Generator().manual_seed(5)
    tgt = torch.rand(T(zt['label']).shape[0], 3, 64, 64, generator=g) * 2 - 1
    prev = [T(zt['prev_label']), T(zt['prev_img'])] if temporal else None
    ld, lg, fake = trainer.train_step(opt, G, D, optG, optD, T(zt['label']), tgt, T(zt['lref']), T(zt['iref']), prev=prev)
    assert torch.isfinite(ld).all() and torch.isfinite(lg).all() and fake.shape == tgt.shape
    assert not torch.equal(G.conv_img.weight.detach(), w0) and not torch.equal(next(D.parameters()).detach(), d0)
