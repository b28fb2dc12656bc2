"""The real drop-in run (INTEGRATION.md option B) at the BENCHED geometry, live against the reference on the same GPU.

The UNMODIFIED reference (baseline/_ref, installed by baseline/install_reference.py; it travels to the GPU box) builds its
own `Vid2VidModel` through `models.models.create_model` twice: once with its own networks (cuDNN/ATen, TF32 off) and once
with `models.networks.define_G/define_D` rebound to `fsv.networks.define_G/define_D`.  The same state_dict (with every
non-spectral conv / SPADE projection re-drawn at trained scale, so that gamma/beta, flow and mask actually move the frame)
and the same synthetic batch go through the reference's own `forward(mode='discriminator'|'generator')`, `LossCollector`,
`FaceRefineModel` and `loss_backward`.  Compared: all D-step and G-step loss tensors, the synthesized frame, flow, mask and
a spread of parameter gradients -- on the exact-fp32 kernels (north-star tolerance 1e-3) and on the default tcgen05 TF32
path (stated tolerance below), at ngf 32 / n_downsample_G 5 (the benchmarked networks).

Nothing here reads /root/reference at run time; without baseline/_ref the tests skip (and say so).
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import refenv   # noqa: E402
import synth    # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'models')),
                                 reason='baseline/_ref missing: run `python baseline/install_reference.py` in the build container')]

from fsvtest import rel_err, l2_err   # noqa: E402


def trained_scale_(sd, seed=0):
    """Re-draw the weights the reference initialises with xavier(gain 0.02) (everything that is not under spectral norm) at
    0.5 / sqrt(fan_in), biases at 0.1 and norm affine parameters around (1, 0): the scale of a trained checkpoint."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if not v.is_floating_point() or k.endswith(('weight_u', 'weight_v', 'running_mean', 'running_var')):
            continue
        if k.endswith('weight_orig'):
            continue                                       # already kaiming-uniform O(fan_in^-1/2) (see ADVICE / layers.init_weights)
        if v.dim() >= 2:
            fan_in = v[0].numel()
            v.copy_(torch.randn(v.shape, generator=g) * (0.5 / fan_in ** 0.5))
        elif k.endswith('bias'):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith('weight'):
            v.copy_(1 + torch.randn(v.shape, generator=g) * 0.1)
    return sd


def _make(kind, H, W, batch, use_fsv, extra=()):
    opt = refenv.parse_opt(kind, H, W, batch, extra=list(extra), gpu=True)
    if use_fsv:
        from fsv import networks
        refenv.patch_networks(networks.define_G, networks.define_D)
    try:
        model, og, od = refenv.create_model(opt)
    finally:
        if use_fsv:
            refenv.unpatch_networks()
    return opt, model, og, od


def _nets(model):
    m = model.module
    return {n: getattr(m, n) for n in ('netG', 'netD', 'netDf', 'netDT') if getattr(m, n, None) is not None}


def _run(opt, model, dl):
    """both modes WITHOUT optimizer steps (so that the two models stay comparable), gradients of each kept."""
    nets = _nets(model)
    for n in nets.values():
        n.zero_grad(set_to_none=True)
    d_losses = model(dl, mode='discriminator')
    sum(x.mean() for x in d_losses).backward()
    gd = {'%s.%s' % (nn, k): p.grad.detach().clone() for nn, net in nets.items() if nn != 'netG' for k, p in net.named_parameters()
          if p.grad is not None}
    for n in nets.values():
        n.zero_grad(set_to_none=True)
    g_losses, generated, _ = model(dl, save_images=True, mode='generator')
    sum(x.mean() for x in g_losses).backward()
    gg = {'netG.%s' % k: p.grad.detach().clone() for k, p in nets['netG'].named_parameters() if p.grad is not None}
    return [x.detach().reshape(()) for x in d_losses], [x.detach().reshape(()) for x in g_losses], generated, gd, gg


GRAD_KEYS_G = ['conv_img.weight', 'up_0.conv_0.weight_orig', 'up_1.bn_0.mlp_gamma2.weight', 'up_2.conv_s.weight_orig', 'up_4.conv_1.weight_orig',
               'fc_spade_0_1.0.weight_orig', 'fc_spade_e_2.4.weight_orig', 'ref_img_down_1.conv.weight_orig', 'ref_label_first.conv.weight_orig',
               'label_embedding.conv_first.0.weight', 'label_embedding.up_4.1.weight', 'img_ref_embedding.down_1.0.weight',
               'flow_network_ref.down_flow.0.0.weight_orig', 'flow_network_ref.res_flow.3.conv_0.weight_orig', 'flow_network_ref.conv_flow.0.weight',
               'flow_network_ref.conv_mask.0.weight']


def _state(model):
    return {n: {k: v.detach().clone() for k, v in net.state_dict().items()} for n, net in _nets(model).items()}


def _restore(model, st):
    for n, net in _nets(model).items():
        net.load_state_dict(st[n])


def _devs(run, ref_run, names):
    """deviation of one run from the fp32 reference run: losses (relative), frame / flow / mask (max-norm relative), warped frame
    (absolute, see below), parameter gradients (relative L2)."""
    d1, g1, gen1, gd1, gg1 = run
    d0, g0, gen0, gd0, gg0 = ref_run
    out = {'loss': {}, 'grad': {}}
    for nm, a, bb in list(zip(names[0], d1, d0)) + list(zip(names[1], g1, g0)):
        out['loss'][nm] = abs(float(a) - float(bb)) / max(abs(float(bb)), 1e-2)
    out['frame'] = rel_err(gen1[0], gen0[0])
    if gen0[3] is not None and gen0[3][0] is not None:
        out['flow'] = rel_err(gen1[3][0], gen0[3][0])
        out['flow_px'] = float((gen1[3][0] - gen0[3][0]).abs().max())
        out['mask'] = rel_err(gen1[4][0], gen0[4][0])
        out['warp_abs'] = float((gen1[2][0] - gen0[2][0]).abs().max())
    for k in ['netG.' + s for s in GRAD_KEYS_G if 'netG.' + s in gg0] + sorted(gd0):
        out['grad'][k] = l2_err((gg1 if k in gg1 else gd1)[k], (gg0 if k in gg0 else gd0)[k])
    return out


def _compare(kind, H, W, batch, use_tc, tol_loss, tol_img, tol_grad, extra=()):
    """Exact path: absolute tolerances.  TF32 path: the same quantities must deviate from the fp32 reference by no more than
    max(tolerance, 3 x the deviation of the REFERENCE's own TF32 mode (cuDNN allow_tf32) from its fp32 mode) -- the yardstick
    for "as accurate as the reference PyTorch path under its default TF32 setting".  The warped frame is compared through the
    flow: the synthetic reference images are white noise, so a sub-pixel flow deviation moves the bilinear sample by
    up to |dflow| x 2 (the image range): |dwarp| <= 2 |dflow_px| + tol is the meaningful bound, and gradients that pass through the
    warp (the flow network's) see its piecewise-constant derivative, hence their looser L2 tolerance."""
    from fsv import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    opt, ref, _, _ = _make(kind, H, W, batch, use_fsv=False, extra=extra)
    old = ops.CONV_USE_TC
    ops.CONV_USE_TC = use_tc
    try:
        opt2, mine, _, _ = _make(kind, H, W, batch, use_fsv=True, extra=extra)
        rn, mn = _nets(ref), _nets(mine)
        assert set(rn) == set(mn)
        for name in rn:
            sd = trained_scale_({k: v.detach().cpu().clone() for k, v in rn[name].state_dict().items()}, seed=len(name))
            rn[name].load_state_dict(sd)
            mn[name].load_state_dict(sd)          # identical keys and shapes: the drop-in contract
            assert type(mn[name]).__module__.startswith('fsv.'), 'the patched factories were not used'
        st = _state(ref)
        b = synth.make(kind, batch, H, W, seed=77)
        dl = refenv.data_list({k: v.cuda() for k, v in b.items()})
        n0 = ops.LAUNCHES[0]
        mine_run = _run(opt2, mine, dl)
        assert ops.LAUNCHES[0] - n0 > 500, 'the fsv kernels did not run'
        ref_run = _run(opt, ref, dl)
        names = ref.module.lossCollector.loss_names_D[:len(ref_run[0])], ref.module.lossCollector.loss_names_G
        dev = _devs(mine_run, ref_run, names)
        yard = None
        if use_tc != 0:
            _restore(ref, st)
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            yard = _devs(_run(opt, ref, dl), ref_run, names)
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False

        def lim(base, y):
            return max(base, 3.0 * y) if yard is not None else base
        for nm, e in dev['loss'].items():
            assert e < lim(tol_loss, yard['loss'][nm] if yard else 0), ('loss', nm, e, yard and yard['loss'][nm])
        assert dev['frame'] < lim(tol_img, yard['frame'] if yard else 0), ('frame', dev['frame'], yard and yard['frame'])
        if 'flow' in dev:
            assert dev['flow'] < lim(tol_img, yard['flow'] if yard else 0), ('flow', dev['flow'], yard and yard['flow'])
            assert dev['mask'] < lim(tol_img, yard['mask'] if yard else 0), ('mask', dev['mask'], yard and yard['mask'])
            assert dev['warp_abs'] < 2.0 * dev['flow_px'] + tol_img, ('warped', dev['warp_abs'], dev['flow_px'])
        worst = ('', 0.0, 0.0)
        for k, e in dev['grad'].items():
            base = tol_grad * (2.0 if 'flow_network' in k else 1.0)
            y = yard['grad'][k] if yard else 0.0
            if e > worst[1]:
                worst = (k, e, y)
            assert e < lim(base, y), ('grad', k, e, y)
        rep = '; '.join('%s %.1e' % kv for kv in dev['loss'].items() if kv[1] > 0)
        print('drop-in %s %dx%d use_tc=%d: loss dev %s; frame %.2e flow %.2e mask %.2e; worst grad L2 %s %.2e' %
              (kind, H, W, use_tc, rep, dev['frame'], dev.get('flow', 0), dev.get('mask', 0), worst[0], worst[1]))
        if yard is not None:
            print('   reference TF32 vs reference fp32 (yardstick): frame %.2e flow %.2e mask %.2e; max loss dev %.1e; that grad %.2e; max grad %.2e' %
                  (yard['frame'], yard.get('flow', 0), yard.get('mask', 0), max(yard['loss'].values()), worst[2], max(yard['grad'].values())))
    finally:
        ops.CONV_USE_TC = old


def test_dropin_face256_exact_fp32_kernels():
    """north-star tolerance: frames within 1e-3 relative fp32 of the reference PyTorch path on identical inputs."""
    _compare('face', 256, 256, 2, use_tc=0, tol_loss=1e-3, tol_img=1e-3, tol_grad=1e-2)


def test_dropin_face256_default_tf32_path():
    """The BENCHED path (tcgen05, TF32 operands, fp32 accumulation) at the benched network geometry.  Stated tolerance for
    TF32 through ~60 layers: frames / flow / mask 5e-3 relative (max-norm), losses 5e-3, parameter gradients 3e-2 relative L2 -- or
    3x the deviation of the reference's own TF32 mode, whichever is larger (see _compare)."""
    _compare('face', 256, 256, 2, use_tc=-1, tol_loss=5e-3, tol_img=5e-3, tol_grad=3e-2)


def test_dropin_pose_add_face_D_exact_fp32_kernels():
    """BASELINE config 3 geometry family (6-channel pose labels, fg-mask D input 20 ch, face discriminator on device-cropped
    regions) at 256x256 through the reference's own LossCollector / FaceRefineModel."""
    _compare('pose', 256, 256, 2, use_tc=0, tol_loss=1e-3, tol_img=1e-3, tol_grad=1e-2)


def test_dropin_pose512_default_tf32_path():
    """The headline workload itself: pose 512x512 --add_face_D, per-GPU batch 2, default (tcgen05) path."""
    _compare('pose', 512, 512, 2, use_tc=-1, tol_loss=5e-3, tol_img=5e-3, tol_grad=3e-2)


def test_dropin_street_exact_fp32_kernels():
    """BASELINE config 4: 20-class one-hot labels (encode_label), no flow branch, D input 46 ch, W = 2H."""
    _compare('street', 128, 256, 2, use_tc=0, tol_loss=1e-3, tol_img=1e-3, tol_grad=1e-2)


def test_reference_train_loop_runs_on_fsv_networks():
    """train.py:55-62 for three iterations (optimizer steps included) with the fsv networks inside the reference's model."""
    opt, mine, og, od = _make('face', 256, 256, 2, use_fsv=True)
    b = synth.make('face', 2, 256, 256, seed=5)
    dl = refenv.data_list({k: v.cuda() for k, v in b.items()})
    first = None
    for it in range(3):
        d, g = refenv.train_iteration(opt, mine, og, od, dl)
        vals = [float(x) for x in list(d) + list(g)]
        assert all(v == v and abs(v) < 1e4 for v in vals), vals
        first = first or vals
    assert vals != first
