#!/usr/bin/env python
"""Generate the golden fixtures that pin the oracle (and, through it, the CUDA path).

Runs ONLY in the build container, where the read-only reference checkout is
mounted at /root/reference.  It imports the reference's own modules (with the
environment shims of SURVEY.md section 8c: an ``apex.parallel.SyncBatchNorm``
alias, a ``dominate`` stub, ``fractions.gcd`` and a device-agnostic ``resample``
-- none of which edits the reference), runs them on CPU in fp32 with fixed seeds
and stores inputs / parameters / outputs / gradients as small ``.npz`` files next
to this script.  Nothing here is imported by the product or by the GPU box.

    python tests/golden/make_golden.py
"""
import fractions
import json
import math
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FSV_REFERENCE', '/root/reference')


def install_shims():
    apex = types.ModuleType('apex')
    par = types.ModuleType('apex.parallel')
    par.SyncBatchNorm = torch.nn.SyncBatchNorm
    apex.parallel = par
    sys.modules['apex'] = apex
    sys.modules['apex.parallel'] = par
    dom = types.ModuleType('dominate')
    dom.tags = types.ModuleType('dominate.tags')
    sys.modules['dominate'] = dom
    sys.modules['dominate.tags'] = dom.tags
    fractions.gcd = math.gcd
    sys.path.insert(0, REF)


def resample_any_device(image, flow):
    # same math as base_network.py:28-37 without the hard .cuda() calls
    b, c, h, w = image.shape
    hor = torch.linspace(-1.0, 1.0, w).view(1, 1, 1, w).expand(b, 1, h, w)
    ver = torch.linspace(-1.0, 1.0, h).view(1, 1, h, 1).expand(b, 1, h, w)
    grid = torch.cat([hor, ver], 1)
    flow = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
    return F.grid_sample(image, (grid + flow).permute(0, 2, 3, 1), mode='bilinear',
                         padding_mode='border', align_corners=True)


TINY = dict(
    n_downsample_G=4, n_downsample_A=2, ngf=4, norm_G='spectralspadesyncbatch', conv_ks=3, embed_ks=1,
    spade_ks=1, spade_combine=True, n_sc_layers=2, add_raw_output_loss=False, adaptive_spade=True,
    no_adaptive_embed=False, adaptive_conv=False, n_adaptive_layers=3, use_label_ref='mul', fineSize=64,
    aspect_ratio=1, n_fc_layers=2, label_nc=0, input_nc=1, output_nc=3, res_for_ref=False,
    netS='encoderdecoder', n_shot=1, lambda_kld=0.0, warp_ref=True, for_face=False, sc_arch='unet',
    norm_F='spectralsyncbatch', nff=4, n_blocks_F=2, n_downsample_F=3, flow_multiplier=20, isTrain=True,
    gpu_ids=[], print_G=False, print_D=False, init_type='xavier', init_variance=0.02, netG='fewshot',
    n_frames_G=2, sep_flow_prev=False, no_sep_warp_embed=False, which_model_netD='multiscale',
    adaptive_D_layers=1, ndf=4, n_layers_D=4, num_D=1, norm_D='spectralinstance', netD_subarch='n_layers',
    gan_mode='hinge', lambda_feat=10.0, lambda_flow=10.0, lambda_mask=10.0, lambda_vgg=10.0,
    lambda_temp=0.0, lambda_face=10.0, no_ganFeat_loss=False, no_vgg_loss=True, no_flow_gt=True,
    dataset_mode='fewshot_face', add_face_D=False, finetune=False, checkpoints_dir='/tmp/fsv_golden',
    name='golden', lr=0.0004, n_frames_per_gpu=1, amp='O0', beta1=0.5, beta2=0.999, no_TTUR=False,
)


def npz_state(prefix, sd):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def synth_face_inputs(gen, B, H, W, K=1):
    def edges(*shape):
        e = (torch.rand(*shape, generator=gen) < 0.05).float()
        return F.max_pool2d(e.view(-1, 1, H, W), 3, 1, 1).view(*shape)
    label = edges(B, 1, H, W)
    lref = edges(B, K, 1, H, W)
    iref = torch.rand(B, K, 3, H, W, generator=gen) * 2 - 1
    tgt = torch.rand(B, 3, H, W, generator=gen) * 2 - 1
    return label, lref, iref, tgt


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print('wrote', path, '%.2f MB' % (os.path.getsize(path) / 1e6))


def main():
    install_shims()
    import models.networks as networks
    import models.networks.generator as refgen
    import models.networks.normalization as refnorm
    import models.networks.base_network as refbase
    import models.loss_collector as reflc
    refgen.resample = resample_any_device
    reflc.resample = resample_any_device

    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234)

    # ---------------------------------------------------------------- generator, training mode
    opt = Namespace(**TINY)
    G = networks.define_G(opt)
    G.train()
    B, H, W = 2, 64, 64
    label, lref, iref, tgt = synth_face_inputs(gen, B, H, W)
    sd0 = {k: v.clone() for k, v in G.state_dict().items()}
    label.requires_grad_(True)
    iref_g = iref.clone().requires_grad_(True)
    out = G(label, lref, iref_g)
    img, flow, mask, raw, warp = out[0], out[1][0], out[2][0], out[3], out[4][0]
    assert raw is None
    r1 = torch.randn(img.shape, generator=gen)
    r2 = torch.randn(flow.shape, generator=gen)
    r3 = torch.randn(mask.shape, generator=gen)
    r4 = torch.randn(warp.shape, generator=gen)
    loss = (img * r1).sum() + 0.05 * (flow * r2).sum() + (mask * r3).sum() + (warp * r4).sum()
    loss.backward()
    grad_names = ['conv_img.weight', 'up_0.conv_0.weight_orig', 'up_0.bn_0.mlp_gamma2.weight',
                  'up_1.conv_s.weight_orig', 'up_3.bn_1.mlp_beta.weight', 'up_4.conv_1.bias',
                  'fc_spade_0_0.0.weight_orig', 'fc_spade_e_1.4.bias', 'fc_spade_s_2.2.weight_orig',
                  'ref_img_first.conv.weight_orig', 'ref_label_down_2.bn.weight', 'ref_img_up_1.conv.bias',
                  'label_embedding.conv_first.0.weight', 'label_embedding.down_2.0.bias',
                  'label_embedding.up_3.1.weight', 'img_ref_embedding.up_0.1.weight',
                  'img_ref_embedding.down_1.0.weight', 'flow_network_ref.down_flow.0.0.weight_orig',
                  'flow_network_ref.res_flow.1.conv_1.weight_orig', 'flow_network_ref.up_flow.4.1.weight',
                  'flow_network_ref.conv_flow.0.weight', 'flow_network_ref.conv_mask.0.bias']
    params = dict(G.named_parameters())
    arrays = dict(opt=json.dumps(TINY), label=label.detach().numpy(), lref=lref.numpy(), iref=iref.numpy(),
                  r1=r1.numpy(), r2=r2.numpy(), r3=r3.numpy(), r4=r4.numpy(),
                  out_img=img.detach().numpy(), out_flow=flow.detach().numpy(), out_mask=mask.detach().numpy(),
                  out_warp=warp.detach().numpy(), loss=np.float64(loss.item()),
                  grad_label=label.grad.numpy(), grad_iref=iref_g.grad.numpy())
    arrays.update(npz_state('sd.', sd0))
    for n in grad_names:
        arrays['grad.' + n] = params[n].grad.numpy()
    sd1 = G.state_dict()
    for k in sd1:   # buffers after one training forward (BN running stats, spectral u/v)
        if k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v', 'num_batches_tracked')):
            if any(s in k for s in ('up_0.', 'ref_img_down_1', 'fc_spade_1_1', 'flow_network_ref.down_flow.2',
                                    'flow_network_ref.res_flow.0', 'ref_label_first', 'up_4.bn_0')):
                arrays['post.' + k] = sd1[k].numpy()
    save('g_face_tiny.npz', **arrays)

    # ---------------------------------------------------------------- generator, eval mode, t=0 then t=1 (weight cache)
    opt_e = Namespace(**dict(TINY, isTrain=False))
    Ge = networks.define_G(opt_e)
    Ge.load_state_dict(sd1)
    Ge.eval()
    with torch.no_grad():
        o0 = Ge(label.detach(), lref, iref, t=0)
        label2, _, _, _ = synth_face_inputs(gen, B, H, W)
        o1 = Ge(label2, lref, iref, t=1)
    arrays = dict(opt=json.dumps(dict(TINY, isTrain=False)), label0=label.detach().numpy(), label1=label2.numpy(),
                  lref=lref.numpy(), iref=iref.numpy(), out_img0=o0[0].numpy(), out_img1=o1[0].numpy(),
                  out_flow1=o1[1][0].numpy(), out_mask1=o1[2][0].numpy())
    arrays.update(npz_state('sd.', {k: v.clone() for k, v in sd1.items()}))
    save('g_face_tiny_eval.npz', **arrays)

    # ---------------------------------------------------------------- generator, temporal phase (warp_prev)
    import util.distributed as refdist
    Gt = networks.define_G(opt)
    Gt.load_state_dict(sd0)
    Gt.init_temporal_network()
    Gt.train()
    sdt0 = {k: v.clone() for k, v in Gt.state_dict().items()}
    prev_label, _, _, prev_img = synth_face_inputs(gen, B, H, W)
    ot = Gt(label.detach(), lref, iref, prev=[prev_label, prev_img])
    lt = (ot[0] * r1).sum() + (ot[4][1] * r4).sum() + (ot[2][1] * r3).sum()
    lt.backward()
    pt = dict(Gt.named_parameters())
    arrays = dict(opt=json.dumps(TINY), label=label.detach().numpy(), lref=lref.numpy(), iref=iref.numpy(),
                  prev_label=prev_label.numpy(), prev_img=prev_img.numpy(), r1=r1.numpy(), r3=r3.numpy(), r4=r4.numpy(),
                  out_img=ot[0].detach().numpy(), out_flow_prev=ot[1][1].detach().numpy(),
                  out_mask_prev=ot[2][1].detach().numpy(), out_warp_prev=ot[4][1].detach().numpy(),
                  loss=np.float64(lt.item()))
    arrays.update(npz_state('sd.', sdt0))
    for n in ['up_0.bn_0.mlp_gamma3.weight', 'img_prev_embedding.up_1.1.weight', 'up_1.conv_0.weight_orig',
              'flow_network_ref.conv_flow.0.weight']:
        arrays['grad.' + n] = pt[n].grad.numpy()
    save('g_face_tiny_temporal.npz', **arrays)

    # ---------------------------------------------------------------- discriminator
    D = networks.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 2, True, gpu_ids=[])
    D.train()
    sdd0 = {k: v.clone() for k, v in D.state_dict().items()}
    xin = torch.randn(4, 8, H, W, generator=gen, requires_grad=True)
    pred = D(xin)
    rs = [[torch.randn(t.shape, generator=gen) for t in p] for p in pred]
    ld = sum((t * r).sum() for p, rr in zip(pred, rs) for t, r in zip(p, rr))
    ld.backward()
    arrays = dict(x=xin.detach().numpy(), grad_x=xin.grad.numpy(), loss=np.float64(ld.item()))
    for i, p in enumerate(pred):
        for j, t in enumerate(p):
            arrays['out.%d.%d' % (i, j)] = t.detach().numpy()
            arrays['r.%d.%d' % (i, j)] = rs[i][j].numpy()
    arrays.update(npz_state('sd.', sdd0))
    for n, p in D.named_parameters():
        arrays['grad.' + n] = p.grad.numpy()
    for k, v in D.state_dict().items():
        if k.endswith(('weight_u', 'weight_v')):
            arrays['post.' + k] = v.numpy()
    save('d_tiny.npz', **arrays)

    # ---------------------------------------------------------------- one train step's losses via the reference LossCollector
    os.makedirs(os.path.join(opt.checkpoints_dir, opt.name), exist_ok=True)
    lc = reflc.LossCollector()
    lc.initialize(opt)
    G2 = networks.define_G(opt)
    G2.load_state_dict(sd0)
    G2.train()
    D1 = networks.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 1, True, gpu_ids=[])
    D1.train()
    sdd1 = {k: v.clone() for k, v in D1.state_dict().items()}
    lab = label.detach()
    tgt_label5, tgt_image5 = lab.unsqueeze(1), tgt.unsqueeze(1)
    nets = (D1, None, None, None)
    # D step (vid2vid_model.py:106-128): G forward under no_grad, then D losses
    with torch.no_grad():
        fake_d = G2(lab, lref, iref)[0]
    data_list = [tgt_label5, [tgt_image5, tgt_image5 * 1], [fake_d.unsqueeze(1), None], lref[:, 0], iref[:, 0]]
    d_losses = lc.compute_GAN_losses(nets, data_list, for_discriminator=True)
    sum(l.mean() for l in d_losses).backward()
    gradD_dstep = {n: p.grad.clone() for n, p in D1.named_parameters()}
    D1.zero_grad()
    # G step (vid2vid_model.py:62-104)
    og = G2(lab, lref, iref)
    fake, flow_g, mask_g, warp_g = og[0], og[1], og[2], og[4]
    data_list = [tgt_label5, [tgt_image5, tgt_image5 * 1], [fake.unsqueeze(1), None], lref[:, 0], iref[:, 0]]
    g_gan, g_feat, _, _ = lc.compute_GAN_losses(nets, data_list, for_discriminator=False)
    rs5 = lambda xs: [x.unsqueeze(1) if x is not None else None for x in xs]  # noqa: E731
    flow5, mask5, warp5 = rs5(flow_g), rs5(mask_g), rs5(warp_g)
    l_flow, l_warp, body = lc.compute_flow_losses(lc.reshape(flow5), lc.reshape(warp5), lc.reshape(tgt_image5),
                                                  [None, None], [None, None], None, tgt_label5, lref[:, 0])
    l_mask = lc.compute_mask_losses(lc.reshape(mask5), fake.unsqueeze(1), lc.reshape(warp5), tgt_label5,
                                    lc.reshape(tgt_image5), None, None, None, body)
    total = g_gan.mean() + g_feat.mean() + l_warp.mean() + l_mask.mean() + l_flow.mean()
    total.backward()
    pg = dict(G2.named_parameters())
    arrays = dict(label=lab.numpy(), lref=lref.numpy(), iref=iref.numpy(), tgt=tgt.numpy(),
                  D_real=d_losses[0].detach().numpy(), D_fake=d_losses[1].detach().numpy(),
                  G_GAN=g_gan.detach().numpy(), G_GAN_Feat=g_feat.detach().numpy(),
                  F_Flow=l_flow.detach().numpy(), F_Warp=l_warp.detach().numpy(), F_Mask=l_mask.detach().numpy(),
                  fake=fake.detach().numpy())
    arrays.update(npz_state('sdD.', sdd1))
    for n in ['discriminator_0.model0.0.weight', 'discriminator_0.model2.0.0.weight_orig',
              'discriminator_0.model4.0.1.weight', 'discriminator_0.model5.0.bias']:
        arrays['gradD.' + n] = gradD_dstep[n].numpy()
    for n in ['conv_img.weight', 'up_2.conv_0.weight_orig', 'fc_spade_0_1.0.weight_orig',
              'flow_network_ref.conv_mask.0.weight', 'img_ref_embedding.conv_first.0.weight']:
        arrays['gradG.' + n] = pg[n].grad.numpy()
    save('step_face_tiny.npz', **arrays)   # generator params = g_face_tiny.npz 'sd.*'

    # ---------------------------------------------------------------- op-level fixtures
    arrays = {}
    for kind, normname in (('batch', 'spectralspadesyncbatch'), ('instance', 'spectralspadeinstance')):
        sp = refnorm.SPADE(8, [4, 6, 4], norm=normname, ks=1, params_free=True)
        for p in sp.parameters():
            torch.nn.init.normal_(p, 0, 0.3, generator=gen)
        sp.train()
        x = torch.randn(3, 8, 10, 12, generator=gen) * 2 + 0.5
        maps = [torch.randn(3, 4, 10, 12, generator=gen), torch.randn(3, 6, 10, 12, generator=gen),
                torch.randn(3, 4, 5, 6, generator=gen)]
        wts = [[[torch.randn(3, 8, 4, 1, 1, generator=gen) * 0.3, torch.randn(3, 8, generator=gen) * 0.3]],
               [[torch.randn(3, 8, 4, 1, 1, generator=gen) * 0.3, torch.randn(3, 8, generator=gen) * 0.3]]]
        y = sp(x, maps, wts)
        pre = 'spade_%s.' % kind
        arrays.update({pre + 'x': x.numpy(), pre + 'm0': maps[0].numpy(), pre + 'm1': maps[1].numpy(),
                       pre + 'm2': maps[2].numpy(), pre + 'wg': wts[0][0][0].numpy(), pre + 'bg': wts[0][0][1].numpy(),
                       pre + 'wb': wts[1][0][0].numpy(), pre + 'bb': wts[1][0][1].numpy(), pre + 'y': y.detach().numpy()})
        arrays.update(npz_state(pre + 'sd.', sp.state_dict()))
    # fixed-weight SPADE, ks=1 with learned map-0 weights too, eval mode (running stats)
    sp = refnorm.SPADE(8, [4], norm='spectralspadesyncbatch', ks=1, params_free=False)
    sp.norm.running_mean.copy_(torch.randn(8, generator=gen) * 0.2)
    sp.norm.running_var.copy_(torch.rand(8, generator=gen) + 0.5)
    sp.eval()
    x = torch.randn(2, 8, 6, 6, generator=gen)
    m = torch.randn(2, 4, 6, 6, generator=gen)
    arrays.update({'spade_eval.x': x.numpy(), 'spade_eval.m0': m.numpy(), 'spade_eval.y': sp(x, m).detach().numpy()})
    arrays.update(npz_state('spade_eval.sd.', sp.state_dict()))
    # warp
    img = torch.randn(2, 3, 9, 13, generator=gen)
    flw = torch.randn(2, 2, 9, 13, generator=gen) * 4
    arrays.update({'warp.img': img.numpy(), 'warp.flow': flw.numpy(), 'warp.out': resample_any_device(img, flw).numpy()})
    # batch_conv + reshape_weight layout (base_network.py:132-167) on an arange tensor
    bn = refbase.BaseNetwork()
    flat = torch.arange(2 * 2 * (5 * 3 + 5), dtype=torch.float32).view(2, -1)
    gb = bn.reshape_weight(flat, [[5, 3, 1, 1]] * 2)
    arrays.update({'reshape.flat': flat.numpy(), 'reshape.gw': gb[0][0].numpy(), 'reshape.gb': gb[0][1].numpy(),
                   'reshape.bw': gb[1][0].numpy(), 'reshape.bb': gb[1][1].numpy()})
    flat_e = torch.arange(2 * (3 * 6 + 3 + 3), dtype=torch.float32).view(2, -1)
    ew = bn.reshape_weight(flat_e[:, :-3], [3, 6, 1, 1])
    arrays.update({'reshape.flat_e': flat_e.numpy(), 'reshape.ew': ew[0].numpy(), 'reshape.eb': ew[1].numpy()})
    xb = torch.randn(2, 6, 4, 4, generator=gen)
    wb = torch.randn(2, 3, 6, 1, 1, generator=gen)
    bb = torch.randn(2, 3, generator=gen)
    arrays.update({'bconv.x': xb.numpy(), 'bconv.w': wb.numpy(), 'bconv.b': bb.numpy(),
                   'bconv.y': refbase.batch_conv(xb, wb, bb).numpy()})
    # reference-feature outer product (generator.py:381-388)
    a = torch.randn(2, 5, 3, 4, generator=gen)
    l = torch.randn(2, 5, 3, 4, generator=gen)
    soft = torch.nn.Softmax(dim=1)(l)
    prod = (a.view(2, 5, 1, 12) * soft.view(2, 1, 5, 12)).sum(3, keepdim=True)
    arrays.update({'outer.a': a.numpy(), 'outer.l': l.numpy(), 'outer.y': prod.numpy()})
    save('ops.npz', **arrays)


def variants():
    """Two more dataset geometries of BASELINE.json (structure only, tiny widths), written to g_variants_tiny.npz:
    'pose'  : 6-channel label, portrait frames H = 2W (fewshot_pose_dataset.py:23-25), warp_ref + spade_combine;
    'street': wide frames W = 2H, 5-channel one-hot-like label, --adaptive_spade only (no warp, no spade_combine)."""
    install_shims()
    import models.networks as networks
    import models.networks.generator as refgen
    import models.loss_collector as reflc
    refgen.resample = resample_any_device
    reflc.resample = resample_any_device
    small = dict(TINY, n_downsample_G=3, n_adaptive_layers=2, n_downsample_F=2, n_blocks_F=1)
    cfgs = {
        'pose': (dict(small, input_nc=6, aspect_ratio=0.5, fineSize=32, dataset_mode='fewshot_pose'), 64, 32),
        'street': (dict(small, input_nc=5, aspect_ratio=2, fineSize=64, warp_ref=False, spade_combine=False,
                        dataset_mode='fewshot_street'), 32, 64),
    }
    arrays = {}
    gen = torch.Generator().manual_seed(4321)
    for name, (cfg, H, W) in cfgs.items():
        torch.manual_seed(7)
        opt = Namespace(**cfg)
        G = networks.define_G(opt)
        G.train()
        B, C = 2, cfg['input_nc']
        label = torch.rand(B, C, H, W, generator=gen) * 2 - 1
        lref = torch.rand(B, 1, C, H, W, generator=gen) * 2 - 1
        iref = torch.rand(B, 1, 3, H, W, generator=gen) * 2 - 1
        sd0 = {k: v.clone() for k, v in G.state_dict().items()}
        label.requires_grad_(True)
        out = G(label, lref, iref)
        r1 = torch.randn(out[0].shape, generator=gen)
        loss = (out[0] * r1).sum()
        if out[1][0] is not None:
            loss = loss + 0.05 * out[1][0].sum() + out[2][0].sum()
        loss.backward()
        pre = name + '.'
        arrays.update({pre + 'opt': json.dumps(cfg), pre + 'label': label.detach().numpy(), pre + 'lref': lref.numpy(),
                       pre + 'iref': iref.numpy(), pre + 'r1': r1.numpy(), pre + 'out_img': out[0].detach().numpy(),
                       pre + 'loss': np.float64(loss.item()), pre + 'grad_label': label.grad.numpy(),
                       pre + 'has_flow': np.int64(out[1][0] is not None)})
        if out[1][0] is not None:
            arrays.update({pre + 'out_flow': out[1][0].detach().numpy(), pre + 'out_mask': out[2][0].detach().numpy()})
        params = dict(G.named_parameters())
        for n in ['conv_img.weight', 'up_0.conv_0.weight_orig', 'fc_spade_0_0.0.weight_orig', 'label_embedding.conv_first.0.weight']:
            arrays[pre + 'grad.' + n] = params[n].grad.numpy()
        arrays.update(npz_state(pre + 'sd.', sd0))
    save('g_variants_tiny.npz', **arrays)


def kshot():
    """K = 2 reference images: the attention module merges the two references' features
    (generator.py:298-316,359-366) and ref_idx picks the one that is warped.  -> g_kshot_tiny.npz"""
    install_shims()
    import models.networks as networks
    import models.networks.generator as refgen
    import models.loss_collector as reflc
    refgen.resample = resample_any_device
    reflc.resample = resample_any_device
    cfg = dict(TINY, n_downsample_G=3, n_adaptive_layers=2, n_downsample_F=2, n_blocks_F=1, n_shot=2, n_downsample_A=2)
    torch.manual_seed(11)
    gen = torch.Generator().manual_seed(99)
    G = networks.define_G(Namespace(**cfg))
    G.train()
    B, H, W, K = 2, 32, 32, 2
    label, lref, iref, _ = synth_face_inputs(gen, B, H, W, K=K)
    sd0 = {k: v.clone() for k, v in G.state_dict().items()}
    label.requires_grad_(True)
    out = G(label, lref, iref)
    r1 = torch.randn(out[0].shape, generator=gen)
    loss = (out[0] * r1).sum() + 0.05 * out[1][0].sum() + out[2][0].sum()
    loss.backward()
    params = dict(G.named_parameters())
    arrays = dict(opt=json.dumps(cfg), label=label.detach().numpy(), lref=lref.numpy(), iref=iref.numpy(), r1=r1.numpy(),
                  out_img=out[0].detach().numpy(), out_flow=out[1][0].detach().numpy(), out_mask=out[2][0].detach().numpy(),
                  out_warp=out[4][0].detach().numpy(), atn_vis=out[7].detach().numpy(), ref_idx=out[8].numpy(),
                  loss=np.float64(loss.item()), grad_label=label.grad.numpy())
    for n in ['atn_key_first.conv.weight_orig', 'atn_query_1.conv.weight_orig', 'up_0.conv_0.weight_orig',
              'fc_spade_0_0.0.weight_orig', 'ref_img_down_2.conv.weight_orig']:
        arrays['grad.' + n] = params[n].grad.numpy()
    arrays.update(npz_state('sd.', sd0))
    save('g_kshot_tiny.npz', **arrays)


def temporal_step():
    """G-step losses of the temporal phase (warp_prev) through the reference LossCollector: the flow / mask terms of BOTH
    branches (reference image and previous frame).  Reuses state and inputs of g_face_tiny_temporal.npz.  -> step_face_tiny_temporal.npz"""
    install_shims()
    import models.networks as networks
    import models.networks.generator as refgen
    import models.loss_collector as reflc
    refgen.resample = resample_any_device
    reflc.resample = resample_any_device
    zt = np.load(os.path.join(HERE, 'g_face_tiny_temporal.npz'))
    opt = Namespace(**TINY)
    G = networks.define_G(opt)
    G.init_temporal_network()
    G.load_state_dict({k[3:]: torch.from_numpy(np.array(zt[k])) for k in zt.files if k.startswith('sd.')})
    G.train()
    torch.manual_seed(3)
    D1 = networks.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 1, True, gpu_ids=[])
    D1.train()
    sdd = {k: v.clone() for k, v in D1.state_dict().items()}
    os.makedirs(os.path.join(opt.checkpoints_dir, opt.name), exist_ok=True)
    lc = reflc.LossCollector()
    lc.initialize(opt)
    t = lambda k: torch.from_numpy(np.array(zt[k]))  # noqa: E731
    label, lref, iref, plabel, pimg = t('label'), t('lref'), t('iref'), t('prev_label'), t('prev_img')
    gen = torch.Generator().manual_seed(77)
    tgt = torch.rand(label.shape[0], 3, label.shape[2], label.shape[3], generator=gen) * 2 - 1
    og = G(label, lref, iref, prev=[plabel, pimg])
    fake, flow_g, mask_g, warp_g = og[0], og[1], og[2], og[4]
    tgt_label5, tgt_image5 = label.unsqueeze(1), tgt.unsqueeze(1)
    nets = (D1, None, None, None)
    data_list = [tgt_label5, [tgt_image5, tgt_image5 * 1], [fake.unsqueeze(1), None], lref[:, 0], iref[:, 0]]
    g_gan, g_feat, _, _ = lc.compute_GAN_losses(nets, data_list, for_discriminator=False)
    rs5 = lambda xs: [x.unsqueeze(1) if x is not None else None for x in xs]  # noqa: E731
    flow5, mask5, warp5 = rs5(flow_g), rs5(mask_g), rs5(warp_g)
    l_flow, l_warp, body = lc.compute_flow_losses(lc.reshape(flow5), lc.reshape(warp5), lc.reshape(tgt_image5),
                                                  [None, None], [None, None], None, tgt_label5, lref[:, 0])
    l_mask = lc.compute_mask_losses(lc.reshape(mask5), fake.unsqueeze(1), lc.reshape(warp5), tgt_label5,
                                    lc.reshape(tgt_image5), None, None, None, body)
    total = g_gan.mean() + g_feat.mean() + l_warp.mean() + l_mask.mean() + l_flow.mean()
    total.backward()
    pg = dict(G.named_parameters())
    arrays = dict(tgt=tgt.numpy(), G_GAN=g_gan.detach().numpy(), G_GAN_Feat=g_feat.detach().numpy(),
                  F_Warp=l_warp.detach().numpy(), F_Mask=l_mask.detach().numpy(), fake=fake.detach().numpy())
    arrays.update(npz_state('sdD.', sdd))
    for n in ['conv_img.weight', 'up_0.bn_0.mlp_gamma3.weight', 'flow_network_ref.conv_mask.0.weight', 'img_prev_embedding.up_1.1.weight']:
        arrays['gradG.' + n] = pg[n].grad.numpy()
    save('step_face_tiny_temporal.npz', **arrays)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'temporal_step':
        temporal_step()
    elif len(sys.argv) > 1 and sys.argv[1] == 'kshot':
        kshot()
    elif len(sys.argv) > 1 and sys.argv[1] == 'variants':
        variants()
    else:
        main()
