"""Network-level CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Functional (no nn.Module): every function takes the flat parameter dict ``sd``
keyed like the reference ``state_dict`` plus an ``opt`` namespace with the
reference's option names (options/base_options.py:54-102), and returns what the
reference module returns.  Scope = SURVEY.md section 8(a) rows a1-a12 for the
north-star configuration: ``--adaptive_spade`` (+``--warp_ref --spade_combine``),
``use_label_ref='mul'``, ``netS='encoderdecoder'``, ``sc_arch='unet'``, K=1
reference, optional temporal ``prev`` inputs, ``netD_subarch='n_layers'``.
Options outside that scope (``adaptive_conv``, ``res_for_ref``, ``lambda_kld``,
K>1 attention, ``concat`` label use) raise NotImplementedError rather than
silently diverge.
"""
import torch
import torch.nn.functional as F

from . import ops
from .ops import lrelu, get_weight, batch_norm, spade, batch_conv, up2


def _norm_kind(norm):
    # normalization.py:32-35
    return 'batch' if 'batch' in norm else 'instance'


def channels(opt):
    """generator.py:26-29."""
    nf_max = min(1024, opt.ngf * (2 ** opt.n_downsample_G))
    return [min(nf_max, opt.ngf * (2 ** i)) for i in range(opt.n_downsample_G + 2)]


def check_scope(opt):
    if getattr(opt, 'adaptive_conv', False) or getattr(opt, 'res_for_ref', False):
        raise NotImplementedError('adaptive_conv / res_for_ref are outside the hot-path scope')
    if getattr(opt, 'lambda_kld', 0) > 0:
        raise NotImplementedError('the kld bottleneck is outside the hot-path scope')
    if 'mul' not in opt.use_label_ref:
        raise NotImplementedError("only use_label_ref='mul' is in scope")


# ------------------------------------------------------------------ building blocks

def sn_conv_bn_lrelu(sd, prefix, x, stride, training):
    """architecture.py:57-69 SPADEConv2d with a non-SPADE norm: spectral conv3x3
    (+bias) -> BatchNorm(affine) -> LeakyReLU(0.2)."""
    w = get_weight(sd, prefix + '.conv', training)
    y = F.conv2d(x, w, sd[prefix + '.conv.bias'], stride=stride, padding=1)
    return lrelu(batch_norm(y, sd, prefix + '.bn', training))


def spade_resblock(sd, prefix, x, maps, norm_kind, training, norm_weights=None, taps=None):
    """architecture.py:92-108 SPADEResnetBlock.forward with SPADE norms
    (generator main branch).  Learned 1x1 shortcut has NO activation (:103)."""
    if norm_weights is None:
        norm_weights = [None] * 3
    learned = (prefix + '.conv_s.weight_orig') in sd or (prefix + '.conv_s.weight') in sd
    if learned:
        xs = spade(x, maps, sd, prefix + '.bn_s', norm_kind, training, norm_weights[2])
        xs = F.conv2d(xs, get_weight(sd, prefix + '.conv_s', training), None)
    else:
        xs = x
    t0 = lrelu(spade(x, maps, sd, prefix + '.bn_0', norm_kind, training, norm_weights[0]))
    t1 = F.conv2d(t0, get_weight(sd, prefix + '.conv_0', training), sd[prefix + '.conv_0.bias'], padding=1)
    t2 = lrelu(spade(t1, maps, sd, prefix + '.bn_1', norm_kind, training, norm_weights[1]))
    dx = F.conv2d(t2, get_weight(sd, prefix + '.conv_1', training), sd[prefix + '.conv_1.bias'], padding=1)
    if taps is not None:     # debugging aid for the tests: expose the intermediates
        taps.update({prefix + '.bn_0': t0, prefix + '.conv_0': t1, prefix + '.bn_1': t2, prefix + '.conv_s': xs})
    return xs + dx


def plain_resblock(sd, prefix, x, training):
    """architecture.py:92-108 with a plain BN norm and fin==fout (flow network,
    generator.py:477-480): x + conv_1(act(bn_1(conv_0(act(bn_0(x))))))."""
    dx = lrelu(batch_norm(x, sd, prefix + '.bn_0', training))
    dx = F.conv2d(dx, get_weight(sd, prefix + '.conv_0', training), sd[prefix + '.conv_0.bias'], padding=1)
    dx = lrelu(batch_norm(dx, sd, prefix + '.bn_1', training))
    dx = F.conv2d(dx, get_weight(sd, prefix + '.conv_1', training), sd[prefix + '.conv_1.bias'], padding=1)
    return x + dx


def label_embedder(sd, prefix, opt, x, unet, params_free_layers=0, weights=None):
    """generator.py:541-572 LabelEmbedder.forward for netS 'encoderdecoder'
    (label pyramid, decoder only output) and 'unet' (skip concatenation)."""
    if x is None:
        return None
    nd = opt.n_downsample_G
    out = [lrelu(F.conv2d(x, sd[prefix + '.conv_first.0.weight'], sd[prefix + '.conv_first.0.bias'], padding=1))]
    for i in range(nd):
        out.append(lrelu(F.conv2d(out[-1], sd['%s.down_%d.0.weight' % (prefix, i)],
                                  sd['%s.down_%d.0.bias' % (prefix, i)], stride=2, padding=1)))
    if not unet:
        out = [out[-1]]
    for i in reversed(range(nd)):
        xi = out[-1]
        if unet and i != nd - 1:
            xi = torch.cat([xi, out[i + 1]], dim=1)
        xi = up2(xi)
        if i >= params_free_layers:
            y = F.conv2d(xi, sd['%s.up_%d.1.weight' % (prefix, i)], sd['%s.up_%d.1.bias' % (prefix, i)], padding=1)
        else:
            y = batch_conv(xi, weights[i])
        out.append(lrelu(y))
    if unet:
        out = out[nd:]
    return out[::-1]


def flow_generator(sd, prefix, opt, label, label_prev, img_prev, training):
    """generator.py:456-504 FlowGenerator."""
    x = torch.cat([label, label_prev, img_prev], dim=1)
    nd = opt.n_downsample_F
    for k in range(nd + 1):
        p = '%s.down_flow.%d' % (prefix, 2 * k)
        x = F.conv2d(x, get_weight(sd, p + '.0', training), None, stride=1 if k == 0 else 2, padding=1)
        x = lrelu(batch_norm(x, sd, p + '.1', training))
    for k in range(opt.n_blocks_F):
        x = plain_resblock(sd, '%s.res_flow.%d' % (prefix, k), x, training)
    for k in range(nd):
        p = '%s.up_flow.%d' % (prefix, 3 * k + 1)
        x = F.conv2d(up2(x), get_weight(sd, p + '.0', training), None, padding=1)
        x = lrelu(batch_norm(x, sd, p + '.1', training))
    flow = F.conv2d(x, sd[prefix + '.conv_flow.0.weight'], sd[prefix + '.conv_flow.0.bias'], padding=1) * opt.flow_multiplier
    mask = torch.sigmoid(F.conv2d(x, sd[prefix + '.conv_mask.0.weight'], sd[prefix + '.conv_mask.0.bias'], padding=1))
    return flow, mask


def hyper_mlp(sd, prefix, x, n_fc_layers, training):
    """generator.py:103-110: spectral Linear -> LReLU (x n_fc_layers) -> spectral Linear."""
    for k in range(n_fc_layers):
        p = '%s.%d' % (prefix, 2 * k)
        x = lrelu(F.linear(x, get_weight(sd, p, training), sd[p + '.bias']))
    p = '%s.%d' % (prefix, 2 * n_fc_layers)
    return F.linear(x, get_weight(sd, p, training), sd[p + '.bias'])


def attention_encode(sd, opt, x, name, training):
    """generator.py:292-296: first conv + n_downsample_A stride-2 convs (spectral conv + BN + LReLU each)."""
    x = sn_conv_bn_lrelu(sd, name + '_first', x, 1, training)
    for i in range(opt.n_downsample_A):
        x = sn_conv_bn_lrelu(sd, '%s_%d' % (name, i), x, 2, training)
    return x


def attention_module(sd, opt, x, label, label_ref, training, attention=None):
    """generator.py:298-316: combine the features of the n = n_shot reference images.  x: (b*n, c, h, w).
    attention (b, n*h*w, h*w) = softmax over the reference positions of key^T query; out = x_flat @ attention."""
    bn, c, h, w = x.shape
    n = opt.n_shot
    b = bn // n
    if attention is None:
        key = attention_encode(sd, opt, label_ref, 'atn_key', training)          # (b*n, c, h, w)
        query = attention_encode(sd, opt, label, 'atn_query', training)          # (b, c, h, w)
        key = key.reshape(b, n, c, -1).permute(0, 1, 3, 2).reshape(b, -1, c)     # b x nhw x c
        query = query.reshape(b, c, -1)                                          # b x c x hw
        attention = torch.softmax(torch.bmm(key, query), dim=1)                  # b x nhw x hw
    xf = x.reshape(b, n, c, h * w).permute(0, 2, 1, 3).reshape(b, c, -1)         # b x c x nhw
    out = torch.bmm(xf, attention).reshape(b, c, h, w)
    atn_vis = attention.reshape(b, n, h * w, h * w).sum(2).reshape(b, n, h, w)
    return out, attention, atn_vis[-1:, 0:1]


def reference_encoding(sd, opt, img_ref, label_ref, training, label=None, n=1):
    """generator.py:341-393 (use_label_ref='mul').  With n = n_shot > 1 the (b*n) reference features are merged
    into b by the attention module after down-sampling level n_downsample_A - 1 (:359-366); returns additionally
    (atn, atn_vis, ref_idx) -- ref_idx = the reference with the largest total attention, used to pick the image
    that gets warped (flow_generation, :425)."""
    nd = opt.n_downsample_G
    x = sn_conv_bn_lrelu(sd, 'ref_img_first', img_ref, 1, training)
    xl = sn_conv_bn_lrelu(sd, 'ref_label_first', label_ref, 1, training)
    atn = atn_vis = ref_idx = None
    for i in range(nd):
        x = sn_conv_bn_lrelu(sd, 'ref_img_down_%d' % i, x, 2, training)
        xl = sn_conv_bn_lrelu(sd, 'ref_label_down_%d' % i, xl, 2, training)
        if n > 1 and i == opt.n_downsample_A - 1:
            x, atn, atn_vis = attention_module(sd, opt, x, label, label_ref, training)
            xl, _, _ = attention_module(sd, opt, xl, None, None, training, atn)
            ref_idx = torch.argmax(atn.reshape(label.shape[0], n, -1).sum(2), dim=1)
    enc_img, enc_lab = [x], [xl]
    for i in reversed(range(nd)):
        enc_img.append(sn_conv_bn_lrelu(sd, 'ref_img_up_%d' % i, enc_img[-1], 1, training))
        enc_lab.append(sn_conv_bn_lrelu(sd, 'ref_label_up_%d' % i, enc_lab[-1], 1, training))
    encoded = [ops.ref_outer_product(a, b) for a, b in zip(enc_img, enc_lab)]
    if n > 1:
        return x, encoded[::-1], atn, atn_vis, ref_idx
    return x, encoded[::-1]


def pick_ref(refs, ref_idx):
    """base_network.py:40-47: refs (b, n, c, h, w) -> (b, c, h, w): reference 0, or the one ref_idx names."""
    if ref_idx is None:
        return refs[:, 0]
    idx = ref_idx.long().view(-1, 1, 1, 1, 1).expand(-1, 1, *refs.shape[2:])
    return refs.gather(1, idx)[:, 0]


def spade_hyper_weights(sd, opt, feat, i, training):
    """generator.py:245-273 get_SPADE_weights + base_network.py:132-174."""
    ch = channels(opt)
    ch_in, ch_out = ch[i], ch[i + 1]
    ch_h = ch[i]                      # generator.py:38-41: ch_hidden[i][0] == ch[i]
    sks, eks = opt.spade_ks, opt.embed_ks
    b, c = feat.shape[0], feat.shape[1]
    x = feat.reshape(b * c, -1)       # base_network.py:169-174
    emb = None
    if not opt.no_adaptive_embed:
        fc_e = hyper_mlp(sd, 'fc_spade_e_%d' % i, x, opt.n_fc_layers, training).reshape(b, -1)
        emb = ops.slice_weight_bias(fc_e[:, :-ch_in], [ch_in, ch_out, eks, eks])  # generator.py:262
    fc_0 = hyper_mlp(sd, 'fc_spade_0_%d' % i, x, opt.n_fc_layers, training).reshape(b, -1)
    fc_1 = hyper_mlp(sd, 'fc_spade_1_%d' % i, x, opt.n_fc_layers, training).reshape(b, -1)
    fc_s = hyper_mlp(sd, 'fc_spade_s_%d' % i, x, opt.n_fc_layers, training).reshape(b, -1)
    w0 = ops.slice_gamma_beta(fc_0, [ch_out, ch_h, sks, sks])
    w1 = ops.slice_gamma_beta(fc_1, [ch_in, ch_h, sks, sks])
    ws = ops.slice_gamma_beta(fc_s, [ch_out, ch_h, sks, sks])
    return emb, [w0, w1, ws]


def generator_forward(sd, opt, label, label_refs, img_refs, prev=(None, None), training=True,
                      temporal=False, cached_weights=None, return_internals=False):
    """generator.py:181-229 FewShotGenerator.forward (a1 of SURVEY.md section 8a).

    Returns the reference 9-tuple (img_final, [flow_ref, flow_prev],
    [mask_ref, mask_prev], img_raw, [warp_ref, warp_prev], mu, logvar, atn_vis,
    ref_idx).  ``temporal`` mirrors ``warp_prev`` (generator.py:155-179);
    ``cached_weights`` mirrors the eval-mode t>0 cache (generator.py:415-418)."""
    check_scope(opt)
    nd = opt.n_downsample_G
    ch = channels(opt)
    kind = _norm_kind(opt.norm_G)
    b, n, _, h, w = img_refs.shape
    img_ref = img_refs.reshape(b * n, -1, h, w)
    label_ref = label_refs.reshape(b * n, -1, h, w)

    # ---- weight generation (generator.py:396-422)
    atn_vis = ref_idx = None
    if cached_weights is None:
        if n > 1:
            assert n == opt.n_shot, 'n_shot must equal the number of reference images'
            x, encoded_ref, _atn, atn_vis, ref_idx = reference_encoding(sd, opt, img_ref, label_ref, training, label, n)
        else:
            x, encoded_ref = reference_encoding(sd, opt, img_ref, label_ref, training)
        emb_w, norm_w = [], []
        for i in range(opt.n_adaptive_layers):
            feat = encoded_ref[min(len(encoded_ref) - 1, i + 1)]
            e, nw = spade_hyper_weights(sd, opt, feat, i, training)
            emb_w.append(e)
            norm_w.append(nw)
    else:
        # eval mode, t>0: the reference still runs the encoder down path for x
        nd_ = opt.n_downsample_G
        x = sn_conv_bn_lrelu(sd, 'ref_img_first', img_ref, 1, training)
        for i in range(nd_):
            x = sn_conv_bn_lrelu(sd, 'ref_img_down_%d' % i, x, 2, training)
        # (label branch output is unused when the cache is hit; BN is in eval mode so no state changes)
        emb_w, norm_w = cached_weights
    adap_embed = opt.adaptive_spade and not opt.no_adaptive_embed
    enc_label = label_embedder(sd, 'label_embedding', opt, label, unet=False,
                               params_free_layers=(opt.n_adaptive_layers if adap_embed else 0),
                               weights=emb_w if adap_embed else None)

    # ---- flow + warp (generator.py:424-445)
    flow, mask, warp, ds = [None, None], [None, None], [None, None], [None, None]
    lref, iref = pick_ref(label_refs, ref_idx), pick_ref(img_refs, ref_idx)     # generator.py:425
    warp_ref = opt.warp_ref and not opt.for_face
    if warp_ref:
        flow[0], mask[0] = flow_generator(sd, 'flow_network_ref', opt, label, lref, iref, training)
        warp[0] = ops.resample(iref, flow[0])[:, :3]
    label_prev, img_prev = prev
    if temporal and label_prev is not None:
        # generator.py:159-166: unless a separate temporal flow net was asked for, flow_network_temp IS
        # flow_network_ref (one module called twice: its spectral u/v and BN running stats advance twice)
        sep_flow = opt.sep_flow_prev or (opt.n_frames_G != 2) or not opt.warp_ref
        flow[1], mask[1] = flow_generator(sd, 'flow_network_temp' if sep_flow else 'flow_network_ref', opt,
                                          label, label_prev, img_prev, training)
        warp[1] = ops.resample(img_prev[:, -3:], flow[1])
    if opt.spade_combine:
        if warp_ref:
            ds[0] = torch.cat([warp[0], mask[0]], dim=1)
        if temporal and label_prev is not None:
            ds[1] = torch.cat([warp[1], mask[1]], dim=1)
        # generator.py:448-454
        emb_ref = label_embedder(sd, 'img_ref_embedding', opt, ds[0], unet=('unet' in opt.sc_arch))
        sep_emb = (not opt.no_sep_warp_embed) or not opt.warp_ref      # generator.py:160
        emb_prev = label_embedder(sd, 'img_prev_embedding' if sep_emb else 'img_ref_embedding', opt, ds[1],
                                  unet=('unet' in opt.sc_arch)) if ds[1] is not None else None
        for i in range(opt.n_sc_layers):
            enc_label[i] = [enc_label[i], emb_ref[i] if emb_ref is not None else None,
                            emb_prev[i] if emb_prev is not None else None]

    # ---- main branch (generator.py:199-207)
    internals = {}
    for i in range(nd, -1, -1):
        nw = norm_w[i] if (opt.adaptive_spade and i < opt.n_adaptive_layers) else None
        x = spade_resblock(sd, 'up_%d' % i, x, enc_label[i], kind, training, nw, taps=internals if return_internals else None)
        if return_internals:
            internals['up_%d' % i] = x
        if i != 0:
            x = up2(x)
    img_raw = torch.tanh(F.conv2d(lrelu(x), sd['conv_img.weight'], sd['conv_img.bias'], padding=1))

    # ---- composite (generator.py:213-227)
    if not opt.spade_combine:
        if warp_ref:
            img_final = img_raw * mask[0] + warp[0] * (1 - mask[0])
        else:
            img_final = img_raw
            if not temporal:
                img_raw = None
        if temporal and label_prev is not None:
            img_final = img_final * mask[1] + warp[1] * (1 - mask[1])
    else:
        img_final = img_raw
        img_raw = None
    out = (img_final, flow, mask, img_raw, warp, None, None, atn_vis, ref_idx)
    if return_internals:
        internals.update(enc_label=enc_label, norm_w=norm_w, emb_w=emb_w)
        return out, internals
    return out


# ------------------------------------------------------------------ discriminator

def nlayer_discriminator(sd, prefix, x, n_layers, training=True):
    """discriminator.py:61-102 NLayerDiscriminator with norm 'spectralinstance'
    (normalization.py:54-88): conv4x4 s2 p2 + LReLU; (n_layers-1) x [sn-conv4x4
    s2 (no bias) + InstanceNorm(affine, eps 0.1) + LReLU]; sn-conv4x4 s1 + IN +
    LReLU; conv4x4 s1 -> 1 channel.  Returns all n_layers+2 feature maps."""
    feats = []
    p = prefix + '.model0.0'
    x = lrelu(F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=2, padding=2))
    feats.append(x)
    for k in range(1, n_layers + 1):
        p = '%s.model%d.0' % (prefix, k)
        stride = 2 if k < n_layers else 1
        wgt = get_weight(sd, p + '.0', training)
        x = F.conv2d(x, wgt, None, stride=stride, padding=2)
        x = lrelu(ops.instance_norm(x, sd[p + '.1.weight'], sd[p + '.1.bias'], eps=0.1))
        feats.append(x)
    p = '%s.model%d.0' % (prefix, n_layers + 1)
    x = F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=1, padding=2)
    feats.append(x)
    return feats


def discriminator_forward(sd, x, n_layers=4, num_D=1, training=True):
    """discriminator.py:49-58 MultiscaleDiscriminator.forward (getIntermFeat)."""
    result = []
    for i in range(num_D):
        result.append(nlayer_discriminator(sd, 'discriminator_%d' % i, x, n_layers, training))
        if i != num_D - 1:
            x = ops.avgpool3s2(x)
    return result


# ------------------------------------------------------------------ one training step's losses

def d_input(tgt_label, fake, real, ref_label, ref_image):
    """loss_collector.py:47-58,104-110: batch = [fake ; real], channels =
    [ref_label, ref_image, tgt_label, image] (concat_ref_for_D)."""
    tgt = torch.cat([fake, real], dim=0)
    tgt = torch.cat([tgt_label.repeat(2, 1, 1, 1), tgt], dim=1)
    ref = torch.cat([ref_label, ref_image], dim=1).repeat(2, 1, 1, 1)
    return torch.cat([ref, tgt], dim=1)


def split_pred(pred):
    """base_model.py:141-147 divide_pred."""
    fake = [[t[:t.shape[0] // 2] for t in p] for p in pred]
    real = [[t[t.shape[0] // 2:] for t in p] for p in pred]
    return fake, real


def masked_l1(a, b, m):
    """loss.py:130-138 MaskedL1Loss."""
    m = m.expand_as(a)
    return (a * m - b * m).abs().mean()


def mask_loss(flow_mask, warped, tgt, lambda_mask):
    """loss_collector.py:164-204 (non-pose branch) compute_mask_loss."""
    conf = torch.clamp(1 - (warped - tgt).abs().sum(dim=1, keepdim=True), 0, 1)
    zero, one = torch.zeros_like(flow_mask), torch.ones_like(flow_mask)
    return (masked_l1(flow_mask, zero, conf) + masked_l1(flow_mask, one, 1 - conf)) * lambda_mask


def generator_losses(sdG, sdD, opt, tgt_label, tgt_image, ref_labels, ref_images, n_layers_D=4, num_D=1, prev=None):
    """vid2vid_model.py:62-104 forward_generator for the single-frame phase with
    ``--no_flow_gt --no_vgg_loss`` on a non-pose dataset: returns dict of the
    non-zero losses (G_GAN, G_GAN_Feat, F_Warp, F_Mask) and the fake image."""
    if prev is None:
        out = generator_forward(sdG, opt, tgt_label, ref_labels, ref_images, training=True)
    else:       # temporal phase: both branches of flow / mask / warp contribute (loss_collector.py:132-136,165-168)
        out = generator_forward(sdG, opt, tgt_label, ref_labels, ref_images, prev=prev, training=True, temporal=True)
    fake, flow, fmask, _, warp = out[0], out[1], out[2], out[3], out[4]
    ref_label, ref_image = ref_labels[:, 0], ref_images[:, 0]
    pred = discriminator_forward(sdD, d_input(tgt_label, fake, tgt_image, ref_label, ref_image), n_layers_D, num_D)
    pf, pr = split_pred(pred)
    # loss_collector.py:66 calls criterionGAN(pred_fake, True) WITHOUT for_discriminator=False, so the
    # generator's GAN term is the discriminator-style hinge -mean(min(D(fake)-1, 0)) (loss.py:72-75).
    losses = {'G_GAN': ops.gan_loss(pf, True, for_discriminator=True),
              'G_GAN_Feat': ops.feat_match_loss(pr, pf, opt.lambda_feat)}
    if flow[0] is not None:
        losses['F_Warp'] = (warp[0] - tgt_image).abs().mean() * opt.lambda_flow   # loss_collector.py:154-162
        losses['F_Mask'] = mask_loss(fmask[0], warp[0], tgt_image, opt.lambda_mask)
    if flow[1] is not None:
        losses['F_Warp'] = losses.get('F_Warp', 0) + (warp[1] - tgt_image).abs().mean() * opt.lambda_flow
        losses['F_Mask'] = losses.get('F_Mask', 0) + mask_loss(fmask[1], warp[1], tgt_image, opt.lambda_mask)
    return losses, fake


def discriminator_losses(sdD, tgt_label, fake, tgt_image, ref_label, ref_image, n_layers_D=4, num_D=1):
    """vid2vid_model.py:106-128 forward_discriminator given the (no-grad) fake."""
    pred = discriminator_forward(sdD, d_input(tgt_label, fake.detach(), tgt_image, ref_label, ref_image), n_layers_D, num_D)
    pf, pr = split_pred(pred)
    return {'D_real': ops.gan_loss(pr, True), 'D_fake': ops.gan_loss(pf, False)}
