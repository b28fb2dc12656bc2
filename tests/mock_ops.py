"""TEST INFRASTRUCTURE ONLY: a torch-on-CPU emulation of the ``fsv.ops`` API (the autograd wrappers around the C ABI).

Purpose: the drop-in modules (``fsv.networks``) are Python host logic around ``fsv.ops``.  With this module patched in
for ``ops`` the *same module code* runs on CPU, so the ``-m "not gpu"`` suite can check the host wiring (which op is
called with which tensors / offsets / flags, state-dict handling, weight cache, temporal phase, K-shot attention)
against the golden fixtures produced by the reference -- without a GPU and without any CPU fallback in the product:
nothing under ``few-shot-vid2vid_b200/`` imports this file, and ``fsv.ops`` itself still rejects CPU tensors.

Every function states the contract of the op it stands in for (layouts are NHWC exactly as in ``fsv.ops``).
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_LRELU, ACT_TANH, ACT_SIGMOID, ACT_RELU = 0, 1, 2, 3, 4
NORM_BATCH, NORM_INSTANCE = 0, 1
CONV_USE_TC = 0
LAUNCHES = [0]
SPECTRAL_EMIT_WT = False


def _act(x, act):
    if act == ACT_LRELU:
        return F.leaky_relu(x, 0.2)
    if act == ACT_TANH:
        return torch.tanh(x)
    if act == ACT_SIGMOID:
        return torch.sigmoid(x)
    if act == ACT_RELU:
        return F.relu(x)
    return x


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------ layout
def pad_channels(c):
    return c if c <= 8 else (c + 31) // 32 * 32


def _padc(x_nhwc):
    c = x_nhwc.shape[3]
    return F.pad(x_nhwc, (0, pad_channels(c) - c))


def to_nhwc(x, pad=True):
    return _padc(_nhwc(x)) if pad else _nhwc(x)


def to_nchw(x):
    return _nchw(x).contiguous()


def nchw_view(x):
    return x.permute(0, 3, 1, 2)


def pack_nhwc(*xs):
    return _padc(_nhwc(torch.cat(xs, dim=1)))


def pack_rows(dst, row0, tensors, coff):
    for t in tensors:
        b, c = t.shape[:2]
        dst[row0:row0 + b, :, :, coff:coff + c] = t.detach().permute(0, 2, 3, 1)
        coff += c
    return coff


def d_input(base, fake, coff):
    b, c = fake.shape[:2]
    x = base.clone()
    return torch.cat([torch.cat([x[:b, :, :, :coff], fake.permute(0, 2, 3, 1), x[:b, :, :, coff + c:]], dim=3), x[b:]], dim=0)


def cat_channels(*xs):
    return torch.cat(xs, dim=3)


def upsample2x(x):
    return _nhwc(F.interpolate(_nchw(x), scale_factor=2, mode='nearest'))


def maxpool2(x):
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def avgpool3s2(x):
    return _nhwc(F.avg_pool2d(_nchw(x), 3, stride=2, padding=1, count_include_pad=False))


# ------------------------------------------------------------------ convs
def to_ohwi(w):
    return w.permute(0, 2, 3, 1).contiguous()


def guard_param(p):
    return p


def conv2d(x, w_ohwi, bias=None, stride=1, pad=0, up=1, act=ACT_NONE, out_scale=1.0, residual=None, use_tc=None,
           in_act=ACT_NONE, wt=None, side_ok=False):
    """y = act(conv(up2?(in_act(x)), w) + bias + residual) * out_scale; w is (Cout, kh, kw, Cin)."""
    xin = _act(_nchw(x), in_act)
    if up == 2:
        xin = F.interpolate(xin, scale_factor=2, mode='nearest')
    y = F.conv2d(xin, w_ohwi.permute(0, 3, 1, 2), bias, stride=stride, padding=pad)
    if residual is not None:
        y = y + _nchw(residual)
    return _nhwc(_act(y, act) * out_scale)


def linear(x2d, w, bias, act=ACT_NONE, wt=None, side_ok=False):
    return _act(F.linear(x2d, w, bias), act)


def batch_conv1x1(x, flat, cout, cin, w_off, b_off, act=ACT_NONE):
    """per-sample 1x1 conv; sample b uses flat[b, w_off : w_off + cout*cin] as (cout, cin) and flat[b, b_off : b_off + cout]."""
    b = x.shape[0]
    w = flat[:, w_off:w_off + cout * cin].reshape(b, cout, cin)
    bias = flat[:, b_off:b_off + cout]
    y = torch.einsum('bhwc,boc->bhwo', x, w) + bias[:, None, None, :]
    return _act(y, act)


def per_sample_matmul(x, flat, cout, cin):
    b = x.shape[0]
    return torch.einsum('bhwc,boc->bhwo', x, flat.reshape(b, cout, cin))


def softmax_channels(x):
    return torch.softmax(x, dim=-1)


def softmax_outer(img, lab):
    """out[b, c1, c2] = sum_hw img[b, hw, c1] * softmax_c(lab)[b, hw, c2]."""
    return torch.einsum('bhwi,bhwj->bij', img, torch.softmax(lab, dim=-1))


# ------------------------------------------------------------------ normalisation
def _norm(x_nchw, mode, training, rm, rv, eps, momentum, weight=None, bias=None):
    if mode == NORM_BATCH:
        return F.batch_norm(x_nchw, rm, rv, weight, bias, training, momentum, eps)
    return F.instance_norm(x_nchw, None, None, weight, bias, True, momentum, eps)


def norm_act(x, weight, bias, running_mean, running_var, mode, training, eps, momentum=0.1, act=ACT_NONE):
    return _nhwc(_act(_norm(_nchw(x), mode, training, running_mean, running_var, eps, momentum, weight, bias), act))


class SpadeFn:
    """fused SPADE contract: x (N, H/up, W/up, C); per map i tensors[5i:5i+5] = (map, wg, bg, wb, bb) with fixed weights
    (C, K) / biases (C) or, when cfg['maps'][i] has 'nstride', per-sample weights at flat[b, wg_off/wb_off : + C*K] and NO
    bias; out = act(chain_i v*(1+gamma_i)+beta_i) starting from the normalised (batch / instance) x, upsampled x``up``."""

    @staticmethod
    def apply(x, running_mean, running_var, cfg, *tensors):
        xn = _nchw(x)
        if cfg['up'] == 2:
            xn = F.interpolate(xn, scale_factor=2, mode='nearest')
        v = _norm(xn, cfg['mode'], cfg['training'], running_mean, running_var, cfg['eps'], cfg.get('momentum', 0.1))
        v = v.permute(0, 2, 3, 1)
        b, c = v.shape[0], v.shape[3]
        for i, mc in enumerate(cfg['maps']):
            m, g0, g1, b0, b1 = tensors[5 * i:5 * i + 5]
            K = mc['K']
            assert m.shape[3] == K and tuple(m.shape[:3]) == tuple(v.shape[:3])
            if 'nstride' in mc:
                wg = g0[:, mc['wg_off']:mc['wg_off'] + c * K].reshape(b, c, K)
                wb = b0[:, mc['wb_off']:mc['wb_off'] + c * K].reshape(b, c, K)
                gamma = torch.einsum('bhwk,bck->bhwc', m, wg)
                beta = torch.einsum('bhwk,bck->bhwc', m, wb)
                assert g1 is None and b1 is None
            else:
                gamma = torch.einsum('bhwk,ck->bhwc', m, g0.reshape(c, K)) + (g1 if g1 is not None else 0)
                beta = torch.einsum('bhwk,ck->bhwc', m, b0.reshape(c, K)) + (b1 if b1 is not None else 0)
            v = v * (1 + gamma) + beta
        return _act(v, cfg.get('act', ACT_NONE)).contiguous()


# ------------------------------------------------------------------ warp
def _resample(img_nhwc, flow_nhwc):
    """base_network.py:13-37 through the oracle's restatement (flow in pixels, bilinear, border, align_corners=True)."""
    from oracle import ops as OO
    return OO.resample(_nchw(img_nhwc), _nchw(flow_nhwc)).permute(0, 2, 3, 1)


def warp_concat(img, flow, mask):
    wp = _resample(img, flow)
    return (torch.cat([wp, mask], dim=3) if mask is not None else wp).contiguous()


def warp_blend(img, flow, mask, raw):
    return (raw * mask + _resample(img, flow) * (1 - mask)).contiguous()


# ------------------------------------------------------------------ spectral norm
def spectral_weight(w_orig, u, v, training, eps=1e-12, want_wt=False):
    """one power iteration in training mode (u, v advanced in place), W / sigma returned as OHWI."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v_new = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
            u_new = F.normalize(torch.mv(wm, v_new), dim=0, eps=eps)
            v.copy_(v_new)
            u.copy_(u_new)
    sigma = torch.dot(u.clone(), torch.mv(wm, v.clone()))
    w = w_orig / sigma
    out = w.permute(0, 2, 3, 1).contiguous() if w.dim() == 4 else w
    return (out, None) if want_wt else out


# ------------------------------------------------------------------ pose label preprocessing + face region (oracle restatements)
def fg_mask(label, ch=2, thr=-1.0):
    assert ch == 2 and thr == -1.0
    from oracle import ops as OO
    return OO.fg_mask(label)


def face_mask_avg15(label, ch=2):
    from oracle import ops as OO
    return OO.face_mask_avg15(label[:, ch])


def part_masks(label, ch=2):
    from oracle import ops as OO
    return OO.part_masks(label[:, ch]).permute(0, 2, 3, 1).contiguous()


def face_bbox(planes, thr, openpose, crop_smaller=0):
    """(B, 4) int32 boxes; face pixels = all planes > thr."""
    from oracle import ops as OO
    f = None
    for t, ch in planes:
        m = t[:, ch % t.shape[1]] > thr
        f = m if f is None else (f & m)
    h, w = f.shape[1:]
    return torch.tensor([OO.face_region(f[i], h, w, openpose, crop_smaller) for i in range(f.shape[0])], dtype=torch.int32)


def crop_resize(image, box, size):
    from oracle import ops as OO
    return OO.crop_face_region(image, [tuple(int(v) for v in b) for b in box], size).permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------ grouped spectral norm (layers.SpectralPlanner)
class SpectralGroup:
    def __init__(self, entries):
        self.entries = entries
        self.n = len(entries)

    def matches(self, entries):
        return len(entries) == self.n and all(a[0] is b[0] for a, b in zip(self.entries, entries))


def spectral_group_weights(group, training, eps, w_origs):
    ws = [spectral_weight(w, u, v, training, eps) for w, u, v, _ in group.entries]
    return ws, [None] * group.n



# ------------------------------------------------------------------ fused generator-step losses: the reference formulas in torch
def _masked_l1(a, b, m):
    m = m.expand_as(a)
    return F.l1_loss(a * m, b * m)


def flow_mask_losses(warp0, mask0, warp1, mask1, tgt, fake, ref_body_warp, body, ref_fg_warp, fg, face_avg, fg_diff):
    """loss_collector.py:131-204 (values before the lambda factors); frame tensors NHWC, tgt NCHW."""
    tg = tgt.permute(0, 2, 3, 1)
    f_warp = tgt.new_zeros(())
    f_mask = tgt.new_zeros(())
    for wp, m in ((warp0, mask0), (warp1, mask1)):
        if wp is None:
            continue
        f_warp = f_warp + F.l1_loss(wp, tg)
        conf = torch.clamp(1 - torch.sum(abs(wp - tg), dim=3, keepdim=True), 0, 1)
        f_mask = f_mask + _masked_l1(m, torch.zeros_like(m), conf) + _masked_l1(m, torch.ones_like(m), 1 - conf)
    body_diff = None
    if ref_body_warp is not None:
        f_warp = f_warp + F.l1_loss(ref_body_warp, body)
        body_diff = torch.sum(abs(ref_body_warp - body), dim=3, keepdim=True)
    if ref_fg_warp is not None:
        f_warp = f_warp + F.l1_loss(ref_fg_warp, fg)
    if face_avg is not None:
        f_mask = f_mask + _masked_l1(mask0, torch.zeros_like(mask0), face_avg)
        if fake is not None:
            f_mask = f_mask + _masked_l1(fake, warp0.detach(), face_avg)
        f_mask = f_mask + _masked_l1(mask0, torch.ones_like(mask0), fg_diff)
        if body_diff is not None:
            f_mask = f_mask + _masked_l1(mask0, torch.ones_like(mask0), body_diff)
    return torch.stack([f_warp, f_mask])


def halves_l1(x):
    b = x.shape[0] // 2
    return F.l1_loss(x[:b], x[b:].detach()).view(1)
