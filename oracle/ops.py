"""Op-level CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line it follows (paths relative to the
reference checkout, NVlabs/few-shot-vid2vid @ 009e23f1).  Tensors are NCHW like
the reference; dtype follows the inputs (run it in float64 for a tight check).
Parameters live in a flat ``dict`` keyed exactly like the reference
``state_dict``; buffers are updated in place when
``training`` is true, as the reference modules do.

The only torch kernels used are dense ``conv2d`` / ``linear`` / matmul on CPU --
everything the hot path adds on top (normalisation, SPADE modulation, warping,
hyper-weight slicing, spectral normalisation) is written out as explicit math.
"""
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2  # models/networks/architecture.py:15-17


def lrelu(x):
    """architecture.py:15-17 ``actvn``."""
    return torch.where(x > 0, x, x * LRELU_SLOPE)


def _normalize(v, eps=1e-12):
    # torch.nn.functional.normalize(v, dim=0, eps): v / max(||v||_2, eps)
    return v / torch.clamp(v.norm(), min=eps)


def spectral_weight(sd, prefix, training):
    """torch.nn.utils.spectral_norm as used at architecture.py:60,81-84,
    generator.py:106-109, normalization.py:64-65: one power iteration per
    forward in training mode (buffers ``weight_u/_v`` advance in place), none in
    eval mode; ``W = W_orig / (u^T W_mat v)``."""
    w = sd[prefix + '.weight_orig']
    u = sd[prefix + '.weight_u']
    v = sd[prefix + '.weight_v']
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v_new = _normalize(wm.t().mv(u))
            u_new = _normalize(wm.mv(v_new))
            v.copy_(v_new)
            u.copy_(u_new)
    sigma = torch.dot(u.detach().clone(), wm.mv(v.detach().clone()))
    return w / sigma


def get_weight(sd, prefix, training):
    """Plain or spectrally normalised weight, whichever the state_dict holds."""
    if prefix + '.weight_orig' in sd:
        return spectral_weight(sd, prefix, training)
    return sd[prefix + '.weight']


def batch_norm(x, sd, prefix, training, eps=1e-5, momentum=0.1):
    """(Sync)BatchNorm2d as instantiated at normalization.py:33,78-80
    (apex SyncBatchNorm == local-statistics BatchNorm under single-process DP,
    SURVEY.md section 2a).  Training: batch mean / biased variance over (N,H,W)
    for the normalisation, running stats updated with the UNBIASED variance and
    momentum 0.1; eval: running stats.  Affine if the dict has ``weight``."""
    if training:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        with torch.no_grad():
            rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
            rm.mul_(1 - momentum).add_(momentum * mean.detach())
            rv.mul_(1 - momentum).add_(momentum * var.detach() * (n / max(n - 1, 1)))
            sd[prefix + '.num_batches_tracked'] += 1
    else:
        mean, var = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    y = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    if prefix + '.weight' in sd:
        y = y * sd[prefix + '.weight'].view(1, -1, 1, 1) + sd[prefix + '.bias'].view(1, -1, 1, 1)
    return y


def instance_norm(x, weight=None, bias=None, eps=0.1):
    """nn.InstanceNorm2d(eps=0.1) as at normalization.py:35,82: per-sample,
    per-channel mean / biased variance over (H,W), no running stats."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    return y


def nearest_resize(x, size):
    """F.interpolate(x, size=size) (mode 'nearest'), normalization.py:42:
    src = floor(dst * in/out)."""
    h_in, w_in = x.shape[2:]
    h_out, w_out = size
    if (h_in, w_in) == (h_out, w_out):
        return x
    ih = torch.clamp((torch.arange(h_out, dtype=torch.float32) * (h_in / h_out)).floor().long(), max=h_in - 1)
    iw = torch.clamp((torch.arange(w_out, dtype=torch.float32) * (w_in / w_out)).floor().long(), max=w_in - 1)
    return x[:, :, ih][:, :, :, iw]


def up2(x):
    """nearest x2 upsample: generator.py:124,207 / nn.Upsample(scale_factor=2)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def batch_conv(x, weight, bias=None, stride=1):
    """base_network.py:56-71: per-sample conv, weight (B,Cout,Cin,k,k), bias
    (B,Cout); ``weight`` may be the [weight, bias] pair."""
    if weight is None:
        return x
    if isinstance(weight, (list, tuple)):
        weight, bias = weight
    pad = weight.shape[-1] // 2
    ys = []
    for i in range(x.shape[0]):
        ys.append(F.conv2d(x[i:i + 1], weight[i], None if bias is None else bias[i],
                           stride=stride, padding=pad))
    return torch.cat(ys, 0)


def spade(x, maps, sd, prefix, norm_kind, training, weights=None):
    """normalization.py:37-52 ``SPADE.forward``.

    out = norm(x); for each non-None map i: m = nearest-resize(map_i) to x's
    size; (gamma, beta) = fixed conv ``mlp_gamma{s}/mlp_beta{s}`` (i>0 or no
    hyper-weights) or per-sample ``batch_conv`` with weights[0][j] / weights[1][j]
    (map 0 with hyper-weights); out = out*(1+gamma)+beta.
    norm_kind: 'batch' (SyncBN affine=False, eps 1e-5) or 'instance' (eps 0.1)."""
    if not isinstance(maps, list):
        maps = [maps]
    if norm_kind == 'batch':
        out = batch_norm(x, sd, prefix + '.norm', training)
    else:
        out = instance_norm(x, eps=0.1)
    for i, m in enumerate(maps):
        if m is None:
            continue
        m = nearest_resize(m, x.shape[2:])
        if weights is None or i != 0:
            s = str(i + 1) if i > 0 else ''
            wg, bg = sd['%s.mlp_gamma%s.weight' % (prefix, s)], sd['%s.mlp_gamma%s.bias' % (prefix, s)]
            wb, bb = sd['%s.mlp_beta%s.weight' % (prefix, s)], sd['%s.mlp_beta%s.bias' % (prefix, s)]
            pad = wg.shape[-1] // 2
            gamma = F.conv2d(m, wg, bg, padding=pad)
            beta = F.conv2d(m, wb, bb, padding=pad)
        else:
            j = min(i, len(weights[0]) - 1)
            gamma = batch_conv(m, weights[0][j])
            beta = batch_conv(m, weights[1][j])
        out = out * (1 + gamma) + beta
    return out


def resample(image, flow):
    """base_network.py:13-37: grid = linspace(-1,1) mesh + flow/((W-1)/2,(H-1)/2);
    bilinear ``grid_sample`` with padding_mode='border', align_corners=True,
    written out as explicit gather math (flow is in pixels, channel 0 = x)."""
    b, c, h, w = image.shape
    dt = image.dtype
    gx = torch.linspace(-1.0, 1.0, w, dtype=dt).view(1, 1, w) + flow[:, 0] / ((w - 1.0) / 2.0)
    gy = torch.linspace(-1.0, 1.0, h, dtype=dt).view(1, h, 1) + flow[:, 1] / ((h - 1.0) / 2.0)
    ix = ((gx + 1) / 2) * (w - 1)
    iy = ((gy + 1) / 2) * (h - 1)
    ix = torch.clamp(ix, 0, w - 1)
    iy = torch.clamp(iy, 0, h - 1)
    x0 = ix.detach().floor()
    y0 = iy.detach().floor()
    wx1 = ix - x0
    wy1 = iy - y0
    wx0, wy0 = 1 - wx1, 1 - wy1
    x0l, y0l = x0.long(), y0.long()
    x1l, y1l = torch.clamp(x0l + 1, max=w - 1), torch.clamp(y0l + 1, max=h - 1)
    # corners past the border carry exactly zero weight (ix<=W-1), so clamping the
    # index is equivalent to grid_sample's in-bounds test.
    flat = image.reshape(b, c, h * w)

    def gather(yy, xx):
        idx = (yy * w + xx).view(b, 1, h * w).expand(b, c, h * w)
        return flat.gather(2, idx).view(b, c, h, w)

    out = (gather(y0l, x0l) * (wy0 * wx0).unsqueeze(1) + gather(y0l, x1l) * (wy0 * wx1).unsqueeze(1) +
           gather(y1l, x0l) * (wy1 * wx0).unsqueeze(1) + gather(y1l, x1l) * (wy1 * wx1).unsqueeze(1))
    return out


def ref_outer_product(img_feat, label_feat):
    """generator.py:381-388: softmax over channels of the label feature, then
    sum_hw img[b,c1,hw]*softmax(label)[b,c2,hw] -> (b,c,c,1).  Restated as a
    batched matmul (no (b,c,c,hw) temporary)."""
    b, c, h, w = img_feat.shape
    lab = label_feat - label_feat.max(dim=1, keepdim=True).values
    e = torch.exp(lab)
    soft = e / e.sum(dim=1, keepdim=True)
    prod = torch.bmm(img_feat.reshape(b, c, h * w), soft.reshape(b, c, h * w).transpose(1, 2))
    return prod.view(b, c, c, 1)


def avgpool3s2(x):
    """nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False),
    discriminator.py:28."""
    return F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)


# --------------------------------------------------------------------------
# hyper-weight slicing (base_network.py:126-174)
# --------------------------------------------------------------------------

def slice_weight_bias(flat, shape):
    """base_network.py:154-167 for a single [weight, bias] pair: flat (b, n) ->
    weight = flat[:, :-shape[0]] viewed (b,*shape), bias = last shape[0] cols."""
    b = flat.shape[0]
    nb = shape[0]
    w = flat[:, :-nb].reshape([b] + list(shape))
    return [w, flat[:, -nb:]]


def slice_gamma_beta(flat, shape):
    """base_network.py:136-152 + 154-167 for weight_size [[co,ci,k,k]]*2: the
    flat row is split [gamma block | beta block], each block
    [co*ci*k*k weights | co biases]."""
    n = shape[0] * shape[1] * shape[2] * shape[3] + shape[0]
    assert flat.shape[1] == 2 * n, (flat.shape, n)
    return [slice_weight_bias(flat[:, :n], shape), slice_weight_bias(flat[:, n:], shape)]


def hinge_loss(pred, target_is_real, for_discriminator=True):
    """loss.py:69-83 hinge branch of GANLoss.loss."""
    if for_discriminator:
        z = torch.zeros_like(pred)
        if target_is_real:
            return -torch.minimum(pred - 1, z).mean()
        return -torch.minimum(-pred - 1, z).mean()
    return -pred.mean()


def gan_loss(preds, target_is_real, for_discriminator=True):
    """loss.py:92-104 GANLoss.__call__ over list[num_D] of list features."""
    loss = 0
    for p in preds:
        if isinstance(p, list):
            p = p[-1]
        loss = loss + hinge_loss(p, target_is_real, for_discriminator).view(1)
    return loss / len(preds)


def feat_match_loss(pred_real, pred_fake, lambda_feat=10.0):
    """loss_collector.py:206-215: L1 between D intermediates of fake and
    (detached) real, averaged over num_D, times lambda_feat."""
    num_d = len(pred_fake)
    loss = 0
    for i in range(num_d):
        for j in range(len(pred_fake[i]) - 1):
            loss = loss + (pred_fake[i][j] - pred_real[i][j].detach()).abs().mean() / num_d
    return loss * lambda_feat


# ---------------------------------------------------------------------------- pose label preprocessing / face region
PART_GROUPS = [[0], [1, 2], [3, 4], [5, 6], [7, 9, 8, 10], [11, 13, 12, 14], [15, 17, 16, 18], [19, 21, 20, 22], [23, 24]]


def fg_mask(label):
    """models/input_process.py:52-61 get_fg_mask for label_nc == 0: channel 2 (DensePose part id, background exactly -1),
    dilated by a 15x15 max filter, thresholded at -1.  label (B, C, H, W) -> (B, 1, H, W) in {0, 1}."""
    m = F.max_pool2d(label[:, 2:3], 15, stride=1, padding=7)
    return (m > -1).to(label.dtype)


def part_masks(part):
    """input_process.py:63-79 get_part_mask: part (B, H, W) in [-1, 1] -> (B, 9, H, W); id = (part/2+0.5)*24 within 0.1 of j."""
    p = (part / 2 + 0.5) * 24
    out = torch.zeros(part.shape[0], len(PART_GROUPS), *part.shape[1:], dtype=part.dtype)
    for i, grp in enumerate(PART_GROUPS):
        for j in grp:
            out[:, i] = torch.maximum(out[:, i], ((p > j - 0.1) & (p < j + 0.1)).to(part.dtype))
    return out


def face_mask(part):
    """input_process.py:81-93 get_face_mask: ids 23 and 24.  part (B, H, W) -> (B, 1, H, W)."""
    p = (part / 2 + 0.5) * 24
    m = torch.zeros_like(part, dtype=torch.bool)
    for j in (23, 24):
        m = m | ((p > j - 0.1) & (p < j + 0.1))
    return m.to(part.dtype).unsqueeze(1)


def face_mask_avg15(part):
    """loss_collector.py:178-179: AvgPool2d(15, padding=7, stride=1) of the face mask (zero padding counted)."""
    return F.avg_pool2d(face_mask(part), 15, stride=1, padding=7)


def face_region(face, h, w, use_openpose, crop_smaller=0):
    """models/face_refiner.py:52-83 get_face_region for ONE sample given its boolean face-pixel map (h, w)."""
    idx = face.nonzero()
    if idx.shape[0]:
        y, x = idx[:, 0], idx[:, 1]
        ys, ye, xs, xe = int(y.min()), int(y.max()), int(x.min()), int(x.max())
        if use_openpose:
            xc, yc = (xs + xe) // 2, (ys * 3 + ye * 2) // 5
            ylen = int((xe - xs) * 2.5)
        else:
            xc, yc = (xs + xe) // 2, (ys + ye) // 2
            ylen = int((ye - ys) * 1.25)
        ylen = xlen = min(w, max(32, ylen))
        yc = max(ylen // 2, min(h - 1 - ylen // 2, yc))
        xc = max(xlen // 2, min(w - 1 - xlen // 2, xc))
    else:
        yc, xc = h // 4, w // 2
        ylen = xlen = h // 32 * 8
    ys, ye, xs, xe = yc - ylen // 2, yc + ylen // 2, xc - xlen // 2, xc + xlen // 2
    return ys + crop_smaller, ye - crop_smaller, xs + crop_smaller, xe - crop_smaller


def face_pixels(label, use_openpose):
    """face_refiner.py:57-62: OpenPose: the LAST three channels all > 0; DensePose: channel 2 > 0.9.  label (B, C, H, W)."""
    if use_openpose:
        return (label[:, -3] > 0) & (label[:, -2] > 0) & (label[:, -1] > 0)
    return label[:, 2] > 0.9


def crop_face_region(image, boxes, size):
    """face_refiner.py:34-38: per sample image[i, -3:, ys:ye, xs:xe] nearest-resized to (size, size)."""
    outs = [F.interpolate(image[i:i + 1, -3:, ys:ye, xs:xe], size=(size, size)) for i, (ys, ye, xs, xe) in enumerate(boxes)]
    return torch.cat(outs)
