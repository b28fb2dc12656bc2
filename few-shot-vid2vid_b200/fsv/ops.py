"""torch.autograd.Function wrappers over the fsv_b200 C ABI.

Activations are fp32 NHWC tensors of shape (N, H, W, C), contiguous.  PyTorch is used for
device memory, streams and autograd bookkeeping only; every forward/backward body is one
or more calls into libfsv_b200.so on the current CUDA stream.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import lib, check, ptr, stream, ConvDesc, SpadeDesc, PtrArray, c_vp

ACT_NONE, ACT_LRELU, ACT_TANH, ACT_SIGMOID, ACT_RELU = _lib.ACT_NONE, _lib.ACT_LRELU, _lib.ACT_TANH, _lib.ACT_SIGMOID, _lib.ACT_RELU
NORM_BATCH, NORM_INSTANCE = _lib.NORM_BATCH, _lib.NORM_INSTANCE

# global switch for the conv path: -1 auto (tcgen05 when eligible), 0 force SIMT, 1 force tcgen05
CONV_USE_TC = -1
# launch counter (bench.py reports it as gpu_launches): number of C-ABI compute calls issued
LAUNCHES = [0]


# Weight-gradient work (conv wgrad + the spectral-norm backward that consumes it) is enqueued on a side stream: nothing downstream
# in the backward pass needs dW, so it overlaps with the data-gradient chain (the critical path) and fills SMs the small kernels
# leave idle.  The fork is `side.wait_stream(main)`; the join is ONE `main.wait_stream(side)` in a callback that the autograd engine
# runs when the backward pass ends (before any optimizer can read .grad).  Everything is capturable into a CUDA graph (the side
# stream becomes a parallel branch of the graph).  FSV_WGRAD_SIDE=0 keeps everything on one stream.
WGRAD_SIDE_STREAM = os.environ.get('FSV_WGRAD_SIDE', '1') != '0'
# backward of the upsample-collapsed convolutions at source resolution (4/9 of the MACs, no full-resolution dx); 0: 3x3 at the upsampled one
UP2_BWD_SOURCE = os.environ.get('FSV_UP2_BWD', '1') != '0'
# FSV_SIDE_LANES > 1: weight gradients are dealt round-robin onto that many side streams.  The side stream is a FIFO of ~12 ms of
# mostly small-grid kernels per iteration (a 512 -> 512 linear's weight gradient is 64 CTAs, a 64 -> 64 one a single CTA), and what is
# still queued when the data-gradient chain ends drains serially (round-2 timeline: a 2.6 ms tail of one-at-a-time launches).  Lane 0
# is the gathering lane: the grouped spectral backward and the gradient buckets run there after waiting for the lanes used so far.
SIDE_LANES = min(4, max(1, int(os.environ.get('FSV_SIDE_LANES', '3'))))
_SIDE, _SIDE_DIRTY = {}, {}           # lane 0 per device; streams that forked work since the last join
_SIDE_EXTRA, _SIDE_USED, _SIDE_RR = {}, {}, {}


def _side_lanes(dev):
    side = _SIDE.get(dev)
    if side is None:
        side = _SIDE[dev] = torch.cuda.Stream(device=dev)
    extra = _SIDE_EXTRA.get(dev)
    if extra is None or len(extra) != SIDE_LANES - 1:
        extra = _SIDE_EXTRA[dev] = [torch.cuda.Stream(device=dev) for _ in range(SIDE_LANES - 1)]
    for li, st in enumerate([side] + extra):
        _LANES.setdefault(st.cuda_stream, 4 + li)      # reduction-ticket lanes 4 .. 7 (per-weight spectral backward runs on lane 0 of these)
    return [side] + extra


def side_fork(*tensors, gather=False):
    """-> a side stream, made to wait for everything enqueued on the current stream so far; ``tensors`` (allocated on the
    current stream, about to be read on the side stream) are registered with the allocator so their memory is not reused early.
    ``gather``: lane 0, which additionally waits for every other lane used since the last join (consumers of side-stream results)."""
    dev = torch.cuda.current_device()
    lanes = _side_lanes(dev)
    used = _SIDE_USED.setdefault(dev, set())
    if gather or len(lanes) == 1:
        li = 0
    else:
        li = _SIDE_RR.get(dev, 0) % len(lanes)
        _SIDE_RR[dev] = li + 1
    side = lanes[li]
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    if gather:
        for j in sorted(used):
            if j != 0:
                side.wait_stream(lanes[j])
    used.add(li)
    for t in tensors:
        if t is not None:
            t.record_stream(side)
    owners = _SIDE_DIRTY.setdefault(dev, [])
    if not owners:
        torch.autograd.Variable._execution_engine.queue_callback(side_join)
    if all(o.cuda_stream != cur.cuda_stream for o in owners):
        owners.append(cur)
    return side


def side_gather_stream():
    """Lane 0 after waiting for the current stream and for the other lanes in use (fsv.parallel: gradient buckets); the caller joins."""
    dev = torch.cuda.current_device()
    lanes = _side_lanes(dev)
    used = _SIDE_USED.setdefault(dev, set())
    lanes[0].wait_stream(torch.cuda.current_stream())
    for j in sorted(used):
        if j != 0:
            lanes[0].wait_stream(lanes[j])
    return lanes[0]


def side_join():
    """Every stream that forked weight-gradient work since the last join -- and the current stream -- waits for the side streams.
    (Making the forking streams wait, not just whatever stream happens to be current in the engine callback, keeps the join correct
    when a backward pass is driven from a non-default auxiliary stream, e.g. the discriminator update of trainer.train_iteration.)"""
    if not _SIDE:
        return                       # no weight-gradient work was ever forked in this process
    dev = torch.cuda.current_device()
    owners = _SIDE_DIRTY.get(dev)
    if owners:
        lanes = _side_lanes(dev)
        used = sorted(_SIDE_USED.get(dev) or {0})
        cur = torch.cuda.current_stream()
        for o in owners:
            for j in used:
                o.wait_stream(lanes[j])
        if all(o.cuda_stream != cur.cuda_stream for o in owners):
            for j in used:
                cur.wait_stream(lanes[j])
        _SIDE_DIRTY[dev] = []
        _SIDE_USED[dev] = set()
        _SIDE_RR[dev] = 0


def on_side_stream():
    dev = torch.cuda.current_device()
    if dev not in _SIDE:
        return False
    cur = torch.cuda.current_stream()
    return cur == _SIDE[dev] or any(cur == s for s in _SIDE_EXTRA.get(dev, ()))


def _c(t):
    if t is None:
        return None
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        raise _lib.FsvError('fsv ops take float32 tensors, got %s' % t.dtype)
    return t if t.is_contiguous() else t.contiguous()


def _off(t, off_floats):
    return c_vp(t.data_ptr() + 4 * int(off_floats))


# Independent branches of a network's forward run on a second stream (the generator's reference-encoder / hyper-network / label-
# embedding branch next to its flow / warp / image-embedding branch): their kernels are small and latency-bound, so the two
# chains overlap almost perfectly.  autograd replays each node's backward on the stream its forward ran on, so the backward
# pass overlaps the same way.  The one-launch reductions keep their tickets in per-lane banks (fsv_set_reduction_lane): the
# branch stream uses lane 1.  FSV_BRANCH_STREAMS=0 keeps the forward on one stream.
BRANCH_STREAMS = os.environ.get('FSV_BRANCH_STREAMS', '1') != '0'
_BRANCH, _LANES = {}, {}
_LANE_FNS = ('fsv_norm_stats', 'fsv_norm_stats_finalize', 'fsv_norm_apply_bwd', 'fsv_norm_apply_bwd2', 'fsv_spade_norm_bwd', 'fsv_spectral_fwd',
             'fsv_spectral_bwd')


def aux_stream(index):
    """Auxiliary stream ``index`` (1 = generator branch, 2 = discriminator step) of the current device; its one-launch reductions use
    ticket lane ``index``."""
    key = (torch.cuda.current_device(), index)
    s2 = _BRANCH.get(key)
    if s2 is None:
        s2 = _BRANCH[key] = torch.cuda.Stream(device=key[0])
        _LANES[s2.cuda_stream] = index
    return s2


def branch_fork(*tensors, index=1):
    """-> auxiliary stream ``index``, made to wait for everything enqueued on the current stream so far; ``tensors`` (allocated on the
    current stream, about to be read on the auxiliary one) are registered with the allocator."""
    s2 = aux_stream(index)
    s2.wait_stream(torch.cuda.current_stream())
    for t in tensors:
        if t is not None:
            t.record_stream(s2)
    return s2


def branch_join(s2, *tensors):
    """current stream waits for the branch; ``tensors`` (allocated on the branch stream, consumed on the current one) are registered
    with the allocator."""
    cur = torch.cuda.current_stream()
    cur.wait_stream(s2)
    for t in tensors:
        if t is not None:
            t.record_stream(cur)


def _call(fn, *args):
    LAUNCHES[0] += 1
    if _LANES and fn.__name__ in _LANE_FNS:
        lib.fsv_set_reduction_lane(_LANES.get(torch.cuda.current_stream().cuda_stream, 0))
    check(fn(*args), fn.__name__)


# --------------------------------------------------------------------------- layout

class ToNHWC(torch.autograd.Function):
    """NCHW (reference boundary layout) -> NHWC, zero-padded to ``cp`` >= C channels (see pad_channels)."""

    @staticmethod
    def forward(ctx, x, cp):
        x = _c(x)
        n, c, h, w = x.shape
        cp = max(cp or c, c)
        y = (torch.empty if cp == c else torch.zeros)((n, h, w, cp), device=x.device, dtype=torch.float32)
        _call(lib.fsv_nchw_to_nhwc, ptr(x), ptr(y), n, c, h, w, cp, 0, stream())
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h, w, cp = dy.shape
        dx = torch.empty((n, ctx.c, h, w), device=dy.device, dtype=torch.float32)
        _call(lib.fsv_nhwc_to_nchw, ptr(dy), ptr(dx), n, ctx.c, h, w, cp, 0, 0, stream())
        return dx, None


class ToNCHW(torch.autograd.Function):
    """NHWC -> contiguous NCHW (for callers that need contiguous reference-layout outputs)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
        _call(lib.fsv_nhwc_to_nchw, ptr(x), ptr(y), n, c, h, w, c, 0, 0, stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, c, h, w = dy.shape
        dx = torch.empty((n, h, w, c), device=dy.device, dtype=torch.float32)
        _call(lib.fsv_nchw_to_nhwc, ptr(dy), ptr(dx), n, c, h, w, c, 0, stream())
        return dx


class PackNHWC(torch.autograd.Function):
    """Several NCHW tensors (same N,H,W) -> one channel-concatenated NHWC tensor, in one pass per input.
    Fuses the reference's torch.cat(dim=1) at the network inputs (generator.py:499; loss_collector.py:47-58)
    with the boundary layout change."""

    @staticmethod
    def forward(ctx, *xs):
        xs = [_c(x) for x in xs]
        n, _, h, w = xs[0].shape
        cs = [x.shape[1] for x in xs]
        ct = pad_channels(sum(cs))                                   # zero channels up to the tcgen05 K block, if wide
        y = (torch.empty if ct == sum(cs) else torch.zeros)((n, h, w, ct), device=xs[0].device, dtype=torch.float32)
        off = 0
        for x, c in zip(xs, cs):
            if tuple(x.shape) != (n, c, h, w):
                raise _lib.FsvError('pack_nhwc: inconsistent input shapes')
            _call(lib.fsv_nchw_to_nhwc, ptr(x), ptr(y), n, c, h, w, ct, off, stream())
            off += c
        ctx.cs = cs
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h, w, ct = dy.shape
        outs, off = [], 0
        for i, c in enumerate(ctx.cs):
            if ctx.needs_input_grad[i]:
                dx = torch.empty((n, c, h, w), device=dy.device, dtype=torch.float32)
                _call(lib.fsv_nhwc_to_nchw, ptr(dy), ptr(dx), n, c, h, w, ct, off, 0, stream())
                outs.append(dx)
            else:
                outs.append(None)
            off += c
        return tuple(outs)


def pack_nhwc(*xs):
    return PackNHWC.apply(*xs)


def to_nhwc(x, pad=True):
    cp = pad_channels(x.shape[1]) if pad else x.shape[1]
    if cp == x.shape[1] and x.dim() == 4:
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous() and not x.is_contiguous():
            return v             # already an NCHW-shaped view of an NHWC buffer (a network output fed to another network): zero-copy
    return ToNHWC.apply(x, cp)


def to_nchw(x):
    return ToNCHW.apply(x)


def nchw_view(x_nhwc):
    """Zero-copy NCHW-shaped (channels_last strided) view of an NHWC tensor."""
    return x_nhwc.permute(0, 3, 1, 2)


class CatC(torch.autograd.Function):
    """Channel concat of NHWC tensors (torch.cat(dim=1) of the reference: generator.py:499,562-563)."""

    @staticmethod
    def forward(ctx, *xs):
        xs = [_c(x) for x in xs]
        n, h, w = xs[0].shape[:3]
        cs = [x.shape[3] for x in xs]
        ctx.cs = cs
        y = torch.empty((n, h, w, sum(cs)), device=xs[0].device, dtype=torch.float32)
        off = 0
        for x, c in zip(xs, cs):
            _call(lib.fsv_copy_channels, ptr(x), c, 0, ptr(y), sum(cs), off, n * h * w, c, 0, stream())
            off += c
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h, w, ct = dy.shape
        outs, off = [], 0
        for i, c in enumerate(ctx.cs):
            if ctx.needs_input_grad[i]:
                dx = torch.empty((n, h, w, c), device=dy.device, dtype=torch.float32)
                _call(lib.fsv_copy_channels, ptr(dy), ct, off, ptr(dx), c, 0, n * h * w, c, 0, stream())
                outs.append(dx)
            else:
                outs.append(None)
            off += c
        return tuple(outs)


def cat_channels(*xs):
    return CatC.apply(*xs)


class Up2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=torch.float32)
        _call(lib.fsv_upsample2x_fwd, ptr(x), ptr(y), n, h, w, c, stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h2, w2, c = dy.shape
        dx = torch.empty((n, h2 // 2, w2 // 2, c), device=dy.device, dtype=torch.float32)
        _call(lib.fsv_upsample2x_bwd, ptr(dy), ptr(dx), n, h2 // 2, w2 // 2, c, stream())
        return dx


def upsample2x(x):
    return Up2.apply(x)


class AvgPool3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        ctx.shape = (n, h, w, c)
        y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), device=x.device, dtype=torch.float32)
        _call(lib.fsv_avgpool3s2_fwd, ptr(x), ptr(y), n, h, w, c, stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        n, h, w, c = ctx.shape
        dx = torch.empty((n, h, w, c), device=dy.device, dtype=torch.float32)
        _call(lib.fsv_avgpool3s2_bwd, ptr(dy), ptr(dx), n, h, w, c, stream())
        return dx


def avgpool3s2(x):
    return AvgPool3s2.apply(x)


class MaxPool2(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) on NHWC (VGG19 feature stack, vgg.py:45-59)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h // 2, w // 2, c), device=x.device, dtype=torch.float32)
        _call(lib.fsv_maxpool2_fwd, ptr(x), ptr(y), n, h, w, c, stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        n, h, w, c = x.shape
        dx = torch.empty_like(x)
        _call(lib.fsv_maxpool2_bwd, ptr(x), ptr(dy), ptr(dx), n, h, w, c, stream())
        return dx


def maxpool2(x):
    return MaxPool2.apply(x)


# --------------------------------------------------------------------------- convolution

def _conv_desc(n, h, w, cin, cout, kh, kw, stride, pad, up=1, act=ACT_NONE, out_scale=1.0, w_nstride=0, b_nstride=0,
               use_tc=None, in_act=ACT_NONE):
    d = ConvDesc()
    d.N, d.H, d.W, d.Cin, d.x_ld, d.x_coff, d.up = n, h, w, cin, cin, 0, up
    d.Cout, d.kh, d.kw, d.stride, d.pad = cout, kh, kw, stride, pad
    d.Ho = (h + 2 * pad - kh) // stride + 1
    d.Wo = (w + 2 * pad - kw) // stride + 1
    d.y_ld, d.y_coff, d.act, d.out_scale = cout, 0, act, out_scale
    d.w_nstride, d.b_nstride, d.res_ld, d.res_coff = w_nstride, b_nstride, cout, 0
    d.in_act = in_act
    d.use_tc = CONV_USE_TC if use_tc is None else use_tc
    return d


_UP2_SEL = {}


def _up2_selector(device):
    """sel[parity][a][r]: which kernel rows r land on source-row offset a for an output row of that parity
    (even rows: {r0} | {r1, r2};  odd rows: {r0, r1} | {r2})."""
    key = str(device)
    if key not in _UP2_SEL:
        _UP2_SEL[key] = torch.tensor([[[1., 0., 0.], [0., 1., 1.]], [[1., 1., 0.], [0., 0., 1.]]], device=device)
    return _UP2_SEL[key]


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class Conv2dFn(torch.autograd.Function):
    """y = act(conv(x, w) + b + residual) * out_scale on NHWC.

    ``wbase`` holds OHWI weights starting ``w_off`` floats in; with ``w_nstride`` != 0 sample n
    uses wbase.flat[w_off + n*w_nstride : ...] (the hyper-network's flat output, zero-copy).
    Same for ``bbase`` / ``b_off`` / ``b_nstride``.  cfg: dict(cout, kh, kw, stride, pad, up,
    act, out_scale, w_off, b_off, w_nstride, b_nstride).
    """

    @staticmethod
    def forward(ctx, x, wbase, bbase, residual, cfg):
        x, wbase, bbase, residual = _c(x), _c(wbase), _c(bbase), _c(residual)
        n, hs, ws, cin = x.shape
        up = cfg.get('up', 1)
        d = _conv_desc(n, hs * up, ws * up, cin, cfg['cout'], cfg['kh'], cfg['kw'], cfg.get('stride', 1), cfg.get('pad', 0),
                       up, cfg.get('act', ACT_NONE), cfg.get('out_scale', 1.0), cfg.get('w_nstride', 0), cfg.get('b_nstride', 0),
                       cfg.get('use_tc'), cfg.get('in_act', ACT_NONE))
        y = torch.empty((n, d.Ho, d.Wo, d.Cout), device=x.device, dtype=torch.float32)
        done = False
        if up == 2 and d.use_tc != 0 and cfg.get('w_off', 0) == 0 and lib.fsv_conv2d_fwd_tc_up2_eligible(ctypes.byref(d)):
            # conv3x3(up2(x)) == four 2x2-tap convs of x (one per output parity) with row/column-summed weights:
            # no 4x intermediate and 4/9 of the MACs.  Weight prep is a tiny parameter-side einsum.
            w4 = torch.empty((d.Cout, 16, d.Cin), device=x.device, dtype=torch.float32)
            _call(lib.fsv_up2_weights, ptr(wbase), ptr(w4), d.Cout, d.Cin, stream())
            _call(lib.fsv_conv2d_fwd_tc_up2, ctypes.byref(d), ptr(x), ptr(w4), None if bbase is None else _off(bbase, cfg.get('b_off', 0)),
                  ptr(residual), ptr(y), stream())
            done = True
        if not done:
            xin, dcall = x, d
            if up == 2 and d.use_tc != 0:
                du = ConvDesc.from_buffer_copy(d)
                du.up = 1
                if lib.fsv_conv2d_tc_eligible(ctypes.byref(du)):
                    xin = torch.empty((n, d.H, d.W, cin), device=x.device, dtype=torch.float32)
                    _call(lib.fsv_upsample2x_fwd, ptr(x), ptr(xin), n, hs, ws, cin, stream())
                    dcall = du
            _call(lib.fsv_conv2d_fwd, ctypes.byref(dcall), ptr(xin), _off(wbase, cfg.get('w_off', 0)),
                  None if bbase is None else _off(bbase, cfg.get('b_off', 0)), ptr(residual), ptr(y), stream())
        ctx.cfg, ctx.d = cfg, d
        ctx.has_b, ctx.has_r = bbase is not None, residual is not None
        ctx.same_base = bbase is not None and bbase.data_ptr() == wbase.data_ptr() and bbase.numel() == wbase.numel()
        need_y = d.act != ACT_NONE
        ctx.save_for_backward(x, wbase, y if need_y else None)
        ctx.bshape = None if bbase is None else tuple(bbase.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wbase, y = ctx.saved_tensors
        cfg, d = ctx.cfg, ctx.d
        dy = _c(dy)
        st = stream()
        if d.act != ACT_NONE or d.out_scale != 1.0:
            g = torch.empty_like(dy)
            _call(lib.fsv_act_bwd, ptr(y if y is not None else dy), ptr(dy), ptr(g), dy.numel(), d.act, d.out_scale, st)
        else:
            g = dy
        dx = dw = db = dres = None
        # backward of conv3x3(up2(x)) at source resolution: 4x4 / stride-2 convolutions of g (see csrc/layout.cu)
        up2_src = (UP2_BWD_SOURCE and d.up == 2 and d.use_tc != 0 and cfg.get('w_off', 0) == 0 and d.w_nstride == 0
                   and d.kh == 3 and d.kw == 3 and d.stride == 1 and d.pad == 1 and d.in_act == ACT_NONE)
        dx_done = False
        if ctx.needs_input_grad[0] and up2_src:
            df = _conv_desc(d.N, d.H, d.W, d.Cout, d.Cin, 4, 4, 2, 1, use_tc=d.use_tc)
            if lib.fsv_conv2d_tc_eligible(ctypes.byref(df)):
                wt = cfg.get('wt')
                if wt is None or wt.numel() != wbase.numel():
                    wt = wbase.reshape(d.Cout, 3, 3, d.Cin).permute(3, 1, 2, 0).contiguous()
                wf = torch.empty((d.Cin, 16, d.Cout), device=dy.device, dtype=torch.float32)
                _call(lib.fsv_up2_dgrad_weights, ptr(wt), ptr(wf), d.Cin, d.Cout, st)
                dx = torch.empty_like(x)
                _call(lib.fsv_conv2d_fwd, ctypes.byref(df), ptr(g), ptr(wf), None, None, ptr(dx), st)
                dx_done = True
        if ctx.needs_input_grad[0] and not dx_done:
            dfull = torch.empty((d.N, d.H, d.W, d.Cin), device=dy.device, dtype=torch.float32)
            done = False
            if d.use_tc != 0 and cfg.get('w_off', 0) == 0:
                dd = ConvDesc.from_buffer_copy(d)
                dd.up = 1
                if lib.fsv_conv2d_dgrad_tc_eligible(ctypes.byref(dd)):
                    # tcgen05 data gradient: the kernel wants the weight with its channel axes swapped, wt[ci][r][s][co]
                    wt = cfg.get('wt')
                    if wt is None or wt.numel() != wbase.numel():
                        wt = wbase.reshape(d.Cout, d.kh, d.kw, d.Cin).permute(3, 1, 2, 0).contiguous()
                    _call(lib.fsv_conv2d_dgrad_tc, ctypes.byref(dd), ptr(g), ptr(wt), ptr(dfull), st)
                    done = True
            if not done:
                dd = ConvDesc.from_buffer_copy(d)
                dd.up = 1
                _call(lib.fsv_conv2d_dgrad, ctypes.byref(dd), ptr(g), _off(wbase, cfg.get('w_off', 0)), ptr(dfull), 0, st)
            if d.up == 2:
                dx = torch.empty_like(x)
                _call(lib.fsv_upsample2x_bwd, ptr(dfull), ptr(dx), d.N, d.H // 2, d.W // 2, d.Cin, st)
            else:
                dx = dfull
            if d.in_act != ACT_NONE:   # lrelu on load: sign(lrelu(x)) == sign(x), so x serves as the 'post-activation' value
                _call(lib.fsv_act_bwd, ptr(x), ptr(dx), ptr(dx), dx.numel(), d.in_act, 1.0, st)
        need_w = ctx.needs_input_grad[1]
        need_b = ctx.has_b and ctx.needs_input_grad[2]
        if need_w or need_b:
            shared = d.w_nstride == 0 and cfg.get('w_off', 0) == 0
            # shared weights whose gradient goes to a spectral-norm backward or straight to the leaf: off the critical path
            use_side = WGRAD_SIDE_STREAM and shared and cfg.get('side_ok', False) and not on_side_stream()
            side = side_fork(x, g) if use_side else None
            with torch.cuda.stream(side) if side is not None else _NullCtx():
                sw = stream()
                dw = torch.empty_like(wbase) if shared else torch.zeros_like(wbase)
                if need_b:
                    db = dw if ctx.same_base else torch.zeros(ctx.bshape, device=dy.device, dtype=torch.float32)
                w_done = False
                if need_w and shared and up2_src:
                    # g in the role of x, x in the role of dy: dw16 (Cin, 4, 4, Cout), folded into dw (Cout, 3, 3, Cin)
                    dg = _conv_desc(d.N, d.H, d.W, d.Cout, d.Cin, 4, 4, 2, 1, use_tc=d.use_tc)
                    if lib.fsv_conv2d_wgrad_tc_eligible(ctypes.byref(dg)):
                        dw16 = torch.empty((d.Cin, 16, d.Cout), device=dy.device, dtype=torch.float32)
                        ws = torch.empty(int(lib.fsv_conv2d_wgrad_tc_workspace(ctypes.byref(dg))) // 4 + 1, device=dy.device, dtype=torch.float32)
                        _call(lib.fsv_conv2d_wgrad_tc, ctypes.byref(dg), ptr(g), ptr(x), ptr(dw16), ptr(ws), 0, sw)
                        _call(lib.fsv_up2_wgrad_fold, ptr(dw16), ptr(dw), d.Cout, d.Cin, 0, sw)
                        w_done = True
                if not w_done and need_w and shared and d.use_tc != 0 and lib.fsv_conv2d_wgrad_tc_eligible(ctypes.byref(d)):
                    ws = torch.empty(int(lib.fsv_conv2d_wgrad_tc_workspace(ctypes.byref(d))) // 4 + 1, device=dy.device, dtype=torch.float32)
                    _call(lib.fsv_conv2d_wgrad_tc, ctypes.byref(d), ptr(x), ptr(g), ptr(dw), ptr(ws), 0, sw)
                    w_done = True
                if not w_done or need_b:
                    _call(lib.fsv_conv2d_wgrad, ctypes.byref(d), ptr(x), ptr(g),
                          _off(dw, cfg.get('w_off', 0)) if (need_w and not w_done) else None,
                          _off(db, cfg.get('b_off', 0)) if need_b else None, 0 if shared else 1, sw)
            if ctx.same_base:
                db = None          # the single flat gradient is returned through wbase
            if not need_w:
                dw = None
        if ctx.has_r and ctx.needs_input_grad[3]:
            dres = g
        return dx, dw, db, dres, None


def conv2d(x, w_ohwi, bias=None, stride=1, pad=0, up=1, act=ACT_NONE, out_scale=1.0, residual=None, use_tc=None,
           in_act=ACT_NONE, wt=None, side_ok=False):
    """``wt``: optional (Cin, kh, kw, Cout) copy of the weight with swapped channel axes (spectral_weight(want_wt=True));
    saves the transposing copy the tcgen05 data gradient would otherwise make."""
    cout, kh, kw, _ = w_ohwi.shape
    # side_ok: the caller guarantees that the weight / bias gradients are consumed only by stream-aware code (the spectral-norm
    # backward of this module, which continues on the side stream) or by the leaf accumulation at the end of backward
    cfg = dict(cout=cout, kh=kh, kw=kw, stride=stride, pad=pad, up=up, act=act, out_scale=out_scale, use_tc=use_tc,
               in_act=in_act, wt=wt, side_ok=side_ok)
    return Conv2dFn.apply(x, w_ohwi, bias, residual, cfg)


def batch_conv1x1(x, flat, cout, cin, w_off, b_off, act=ACT_NONE):
    """Per-sample 1x1 conv with weights living inside ``flat`` (B, L): base_network.py:56-71."""
    L = flat.shape[1]
    cfg = dict(cout=cout, kh=1, kw=1, stride=1, pad=0, up=1, act=act, w_off=w_off, b_off=b_off, w_nstride=L, b_nstride=L, use_tc=0)
    assert x.shape[3] == cin
    return Conv2dFn.apply(x, flat, flat, None, cfg)


def linear(x2d, w, bias, act=ACT_NONE, wt=None, side_ok=False):
    """F.linear on (rows, K) with w (out, K): a 1x1 conv over a rows x 1 'image'."""
    rows, k = x2d.shape
    wcols = 32 if rows % 32 == 0 else 1      # a 1x1 conv does not care how the rows are arranged as an image; 32-wide rows
    y = conv2d(x2d.reshape(1, rows // wcols, wcols, k), w.reshape(w.shape[0], 1, 1, k), bias, act=act, wt=wt, side_ok=side_ok)   # suit the TMA boxes
    return y.reshape(rows, w.shape[0])


class OhwiFn(torch.autograd.Function):
    """(Cout, Cin, kh, kw) parameter -> OHWI copy for the conv kernels.  Backward re-lays the (side-stream produced) OHWI weight
    gradient out on the side stream, so the leaf receives a contiguous tensor and its accumulation launches nothing."""

    @staticmethod
    def forward(ctx, w):
        return w.permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def backward(ctx, dw):
        side = side_fork(dw, gather=True) if (WGRAD_SIDE_STREAM and dw.is_cuda and not on_side_stream()) else None
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            return dw.permute(0, 3, 1, 2).contiguous()


def to_ohwi(w):
    return OhwiFn.apply(w)


def guard_param(p):
    """Stream safety net for the side-stream weight gradients: a leaf that already holds a gradient when another one arrives
    (a module called twice in one pass, gradient accumulation over several backward calls, .grad bound to a flat bucket) gets an
    in-place add on the main stream -- which must then first wait for the side stream.  The common case (.grad is None: the
    engine just adopts the tensor) costs nothing."""
    if getattr(p, '_fsv_guarded', False) or not WGRAD_SIDE_STREAM or not p.requires_grad:
        return p          # (a frozen parameter -- the VGG feature stack -- never receives a gradient)

    def hook(grad, p=p):
        if p.grad is not None and grad.is_cuda:
            side_join()
        return grad
    p.register_hook(hook)
    p._fsv_guarded = True
    return p


# --------------------------------------------------------------------------- spectral normalisation of weights

# emit the channel-swapped copy of a spectral weight for the tcgen05 data gradient from the spectral kernel itself
SPECTRAL_EMIT_WT = True
# the spectral backward of a whole group in two launches (fsv_spectral_group_bwd); 0 = one pair of launches per weight
GROUP_SPECTRAL_BWD = os.environ.get('FSV_GROUP_SPECTRAL_BWD', '1') != '0'


class SpectralWeightFn(torch.autograd.Function):
    """torch.nn.utils.spectral_norm's weight computation (one power iteration in training mode, ``weight_u/_v`` advanced
    in place) fused with the OIHW -> OHWI repack: weight_orig (R, Cin[, kh, kw]) -> W / sigma as (R[, kh, kw], Cin).
    With ``want_wt`` a second, non-differentiable output holds the same weight as (Cin[, kh, kw], R)."""

    @staticmethod
    def forward(ctx, w_orig, u, v, training, eps, want_wt):
        w = _c(w_orig)
        _lib.require_cuda(u, v)
        R, cin = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
        K = cin * taps
        out = torch.empty((R, w.shape[2], w.shape[3], cin) if w.dim() == 4 else (R, cin), device=w.device, dtype=torch.float32)
        wt = None
        if want_wt:
            wt = torch.empty((cin, w.shape[2], w.shape[3], R) if w.dim() == 4 else (cin, R), device=w.device, dtype=torch.float32)
        uvs = torch.empty(K + R + 1, device=w.device, dtype=torch.float32)
        work = torch.empty(lib.fsv_spectral_workspace(R, K) // 4, device=w.device, dtype=torch.float32)
        _call(lib.fsv_spectral_fwd, ptr(w), ptr(u), ptr(v), R, cin, taps, 1 if training else 0, float(eps), ptr(out), ptr(wt),
              ptr(uvs), ptr(work), stream())
        ctx.save_for_backward(out, uvs)
        ctx.dims = (R, cin, taps, tuple(w_orig.shape))
        if want_wt:
            ctx.mark_non_differentiable(wt)
            return out, wt
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        out, uvs = ctx.saved_tensors
        R, cin, taps, shape = ctx.dims
        side = side_fork(out, uvs, dout, gather=True) if (WGRAD_SIDE_STREAM and not on_side_stream()) else None   # dout usually comes from a side-stream wgrad
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            dout = _c(dout)
            dw = torch.empty(shape, device=dout.device, dtype=torch.float32)
            work = torch.empty(lib.fsv_spectral_workspace(R, cin * taps) // 4, device=dout.device, dtype=torch.float32)
            _call(lib.fsv_spectral_bwd, ptr(dout), ptr(out), ptr(uvs), R, cin, taps, ptr(dw), ptr(work), stream())
        return dw, None, None, None, None, None


def spectral_weight(w_orig, u, v, training, eps=1e-12, want_wt=False):
    """-> W_sn (OHWI), or (W_sn, wt) with ``want_wt`` (wt = None when SPECTRAL_EMIT_WT is off)."""
    if want_wt and SPECTRAL_EMIT_WT and torch.is_grad_enabled():
        return SpectralWeightFn.apply(w_orig, u, v, training, eps, True)
    out = SpectralWeightFn.apply(w_orig, u, v, training, eps, False)
    return (out, None) if want_wt else out


class SpectralGroup:
    """Plan for computing the spectral weights of a fixed list of modules in three launches (fsv_spectral_group_fwd).
    ``entries`` = [(w_orig, u, v, want_wt)]; the device tables hold the parameters' / buffers' addresses, so the plan is valid as
    long as those tensors are not re-allocated (``matches`` checks)."""

    def __init__(self, entries):
        import numpy as np
        n = len(entries)
        self.n = n
        self.dev = entries[0][0].device
        self.ptrs = [(w.data_ptr(), u.data_ptr(), v.data_ptr()) for w, u, v, _ in entries]
        self.shapes = [tuple(w.shape) for w, _, _, _ in entries]
        items = (_lib.SnItem * n)()
        for it, (w, u, v, want) in zip(items, entries):
            _lib.require_cuda(w, u, v)
            if not w.is_contiguous() or w.dtype != torch.float32:
                raise _lib.FsvError('spectral group: weight_orig must be contiguous float32')
            it.w_orig, it.u, it.v = w.data_ptr(), u.data_ptr(), v.data_ptr()
            it.R, it.Cin = w.shape[0], w.shape[1]
            it.taps = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
            it.want_wt = 1 if want else 0
        self.totals = (_lib.c_ll * 9)()
        check(lib.fsv_spectral_group_plan(ctypes.byref(items), n, ctypes.byref(self.totals)), 'fsv_spectral_group_plan')
        self.items = [dict(R=it.R, Cin=it.Cin, taps=it.taps, K=it.K, want_wt=bool(it.want_wt), out_off=it.out_off, wt_off=it.wt_off,
                           uvs_off=it.uvs_off) for it in items]
        maps = []
        for nb in ('blk1', 'blk2', 'blk3'):
            cnt = {'blk1': [it.nchunks * it.rs for it in items], 'blk2': [it.nblk2 for it in items], 'blk3': [it.nblk3 for it in items]}[nb]
            idx = np.repeat(np.arange(n, dtype=np.int32), cnt)
            loc = np.concatenate([np.arange(c, dtype=np.int32) for c in cnt])
            maps.append(np.stack([idx, loc], axis=1))
        self.map_dev = torch.from_numpy(np.ascontiguousarray(np.concatenate(maps, axis=0))).to(self.dev)
        bmaps = []
        for cnt in ([it.nb1 for it in items], [it.nb2 for it in items]):
            idx = np.repeat(np.arange(n, dtype=np.int32), cnt)
            loc = np.concatenate([np.arange(c, dtype=np.int32) for c in cnt])
            bmaps.append(np.stack([idx, loc], axis=1))
        self.map_bwd_dev = torch.from_numpy(np.ascontiguousarray(np.concatenate(bmaps, axis=0))).to(self.dev)
        self.totals_bwd = (_lib.c_ll * 3)(self.totals[6], self.totals[7], self.totals[8])
        # gradient-pointer tables of the grouped backward: [0] for eager calls, one more per CUDA-graph capture (see optim.Adam._table)
        self.ptr_staging = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(4)]
        self.ptr_dev = [torch.empty(n, dtype=torch.int64, device=self.dev) for _ in range(4)]
        self.ptr_used = 1
        self.ptr_event = None
        self.items_dev = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(self.dev)
        self.tickets = torch.zeros(int(self.totals[2]) + 1, device=self.dev, dtype=torch.int32)
        torch.cuda.current_stream().synchronize()

    def matches(self, entries):
        return len(entries) == self.n and all(p == (w.data_ptr(), u.data_ptr(), v.data_ptr()) for p, (w, u, v, _) in zip(self.ptrs, entries))


class GroupSpectralFn(torch.autograd.Function):
    """W_sn (OHWI) [+ channel-swapped copies] of every weight of a SpectralGroup: forward = 3 launches for the whole group; backward =
    the per-weight fsv_spectral_bwd for the weights that received a gradient."""

    @staticmethod
    def forward(ctx, group, training, eps, emit_wt, *w_origs):
        out = torch.empty(int(group.totals[0]) + 4, device=group.dev, dtype=torch.float32)
        work = torch.empty(int(group.totals[1]) + 4, device=group.dev, dtype=torch.float32)
        _call(lib.fsv_spectral_group_fwd, ptr(group.items_dev), ptr(group.map_dev), ctypes.byref(group.totals), 1 if training else 0, float(eps),
              1 if emit_wt else 0, ptr(out), ptr(work), ptr(group.tickets), stream())
        ws, wts = [], []
        for it, shp in zip(group.items, group.shapes):
            R, Cin, K = it['R'], it['Cin'], it['K']
            oshape = (R, shp[2], shp[3], Cin) if len(shp) == 4 else (R, Cin)
            ws.append(out[it['out_off']:it['out_off'] + R * K].view(oshape))
            if emit_wt and it['want_wt']:
                wts.append(out[it['wt_off']:it['wt_off'] + R * K].view((Cin, shp[2], shp[3], R) if len(shp) == 4 else (Cin, R)))
            else:
                wts.append(None)
        ctx.group, ctx.arena = group, out
        ctx.set_materialize_grads(False)          # a weight whose W_sn was not used gets NO gradient (None), like an unused leaf
        real_wts = [t for t in wts if t is not None]
        ctx.mark_non_differentiable(*real_wts)
        ctx.n = len(ws)
        return tuple(ws) + tuple(real_wts)

    @staticmethod
    def backward(ctx, *grads):
        group, out = ctx.group, ctx.arena
        side = side_fork(out, *grads, gather=True) if (WGRAD_SIDE_STREAM and not on_side_stream()) else None
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            st = stream()
            gs = [None if g is None else _c(g) for g in grads[:group.n]]
            if not GROUP_SPECTRAL_BWD:
                res = []
                for g, it, shp in zip(gs, group.items, group.shapes):
                    if g is None:
                        res.append(None)
                        continue
                    R, Cin, K, taps = it['R'], it['Cin'], it['K'], it['taps']
                    dw = torch.empty(shp, device=g.device, dtype=torch.float32)
                    work = torch.empty(lib.fsv_spectral_workspace(R, K) // 4, device=g.device, dtype=torch.float32)
                    _call(lib.fsv_spectral_bwd, ptr(g), _off(out, it['out_off']), _off(out, it['uvs_off']), R, Cin, taps, ptr(dw), ptr(work), st)
                    res.append(dw)
                return (None, None, None, None) + tuple(res)
            # two launches for the whole group: the dW_sn addresses of this call go to the device through a pinned staging buffer (eager
            # calls reuse slot 0; a call made during CUDA-graph capture takes a slot of its own, which the graph's copy node keeps reading)
            slot = 0
            if torch.cuda.is_current_stream_capturing():
                if group.ptr_used >= len(group.ptr_staging):
                    raise _lib.FsvError('spectral group: too many CUDA-graph captures of one plan')
                slot = group.ptr_used
                group.ptr_used += 1
            host = group.ptr_staging[slot]
            if slot == 0 and group.ptr_event is not None:
                group.ptr_event.synchronize()          # the previous eager call's copy out of this staging buffer has long finished; make it certain
            for i, g in enumerate(gs):
                host[i] = 0 if g is None else g.data_ptr()
            group.ptr_dev[slot].copy_(host, non_blocking=True)
            if slot == 0:
                group.ptr_event = torch.cuda.Event()
                group.ptr_event.record()
            dw_arena = torch.empty(int(group.totals[0]) + 4, device=out.device, dtype=torch.float32)
            work = torch.empty(int(group.totals[6]) + 4, device=out.device, dtype=torch.float32)
            _call(lib.fsv_spectral_group_bwd, ptr(group.items_dev), ptr(group.map_bwd_dev), ctypes.byref(group.totals_bwd), ptr(group.ptr_dev[slot]), ptr(out),
                  ptr(dw_arena), ptr(work), ptr(group.tickets), st)
            res = [None if g is None else dw_arena[it['out_off']:it['out_off'] + it['R'] * it['K']].view(shp)
                   for g, it, shp in zip(gs, group.items, group.shapes)]
            for g in gs:
                if g is not None and side is not None:
                    g.record_stream(side)
        return (None, None, None, None) + tuple(res)


def spectral_group_weights(group, training, eps, w_origs):
    """-> ([W_sn_i], [wt_i or None]) for the group's modules."""
    emit = torch.is_grad_enabled() and SPECTRAL_EMIT_WT
    outs = GroupSpectralFn.apply(group, training, eps, emit, *w_origs)
    n = group.n
    ws, rest = list(outs[:n]), list(outs[n:])
    wts = [(rest.pop(0) if (emit and it['want_wt']) else None) for it in group.items]
    return ws, wts


# --------------------------------------------------------------------------- normalisation (+ activation)

def _stats(x, n, hw, c, mode, training, running_mean, running_var, eps, momentum, unbias_mul=1):
    dev = x.device
    groups = n if mode == NORM_INSTANCE else 1
    mean = torch.empty(groups * c, device=dev, dtype=torch.float32)
    rstd = torch.empty(groups * c, device=dev, dtype=torch.float32)
    st = stream()
    if training or mode == NORM_INSTANCE:
        rpg = hw if mode == NORM_INSTANCE else n * hw
        work = torch.empty(int(lib.fsv_norm_work_doubles(groups, c, rpg)), device=dev, dtype=torch.float64)
        upd = 1 if (mode == NORM_BATCH and running_mean is not None) else 0
        _call(lib.fsv_norm_stats_finalize, ptr(x), n, hw, c, c, 0, mode, float(unbias_mul), eps, momentum,
              ptr(running_mean) if upd else None, ptr(running_var) if upd else None, upd, ptr(mean), ptr(rstd), ptr(work), st)
    else:
        _call(lib.fsv_norm_from_running, ptr(running_mean), ptr(running_var), c, eps, ptr(mean), ptr(rstd), st)
    return mean, rstd


NORM_BWD_RECOMPUTE = os.environ.get('FSV_NORM_BWD2', '1') != '0'      # 0: keep y for the backward (fsv_norm_apply_bwd)


class NormActFn(torch.autograd.Function):
    """y = act(norm(x) * weight + bias): BatchNorm (local statistics) or InstanceNorm, + activation."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, cfg):
        x, weight, bias = _c(x), _c(weight), _c(bias)
        n, h, w, c = x.shape
        mode, training = cfg['mode'], cfg['training']
        mean, rstd = _stats(x, n, h * w, c, mode, training, running_mean, running_var, cfg['eps'], cfg.get('momentum', 0.1))
        y = torch.empty_like(x)
        _call(lib.fsv_norm_apply_fwd, ptr(x), ptr(mean), ptr(rstd), ptr(weight), ptr(bias), ptr(y), n, h * w, c, mode,
              cfg.get('act', ACT_NONE), stream())
        ctx.cfg = cfg
        # the backward recomputes a sign-type activation's derivative from x (fsv_norm_apply_bwd2): y is not kept for it
        recompute = NORM_BWD_RECOMPUTE and cfg.get('act', ACT_NONE) in (ACT_NONE, ACT_LRELU, ACT_RELU)
        ctx.save_for_backward(x, None if recompute else y, mean, rstd, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, weight, bias = ctx.saved_tensors
        cfg = ctx.cfg
        dy = _c(dy)
        n, h, w, c = x.shape
        mode = cfg['mode']
        groups = n if mode == NORM_INSTANCE else 1
        batch_stats = 1 if (cfg['training'] or mode == NORM_INSTANCE) else 0
        dx = torch.empty_like(x)
        dwt = torch.empty(c, device=x.device, dtype=torch.float32) if weight is not None else None
        dbs = torch.empty(c, device=x.device, dtype=torch.float32) if weight is not None else None
        rpg = h * w if mode == NORM_INSTANCE else n * h * w
        scratch = torch.empty(2 * groups * c + int(lib.fsv_norm_work_doubles(groups, c, rpg)), device=x.device, dtype=torch.float64)
        if y is None:
            _call(lib.fsv_norm_apply_bwd2, ptr(x), ptr(dy), ptr(mean), ptr(rstd), ptr(weight), ptr(bias), ptr(dx), ptr(dwt), ptr(dbs),
                  ptr(scratch), n, h * w, c, mode, cfg.get('act', ACT_NONE), batch_stats, stream())
        else:
            _call(lib.fsv_norm_apply_bwd, ptr(x), ptr(y), ptr(dy), ptr(mean), ptr(rstd), ptr(weight), ptr(dx), ptr(dwt), ptr(dbs),
                  ptr(scratch), n, h * w, c, mode, cfg.get('act', ACT_NONE), batch_stats, stream())
        return dx, dwt, dbs, None, None, None


def norm_act(x, weight, bias, running_mean, running_var, mode, training, eps, momentum=0.1, act=ACT_NONE):
    cfg = dict(mode=mode, training=training, eps=eps, momentum=momentum, act=act)
    return NormActFn.apply(x, weight, bias, running_mean, running_var, cfg)


# --------------------------------------------------------------------------- fused SPADE

class SpadeFn(torch.autograd.Function):
    """Fused SPADE: normalise (+ upsample-on-load) -> per-map 1x1 gamma/beta -> modulate -> activation.

    forward(ctx, x, running_mean, running_var, cfg, *tensors) with, per map i,
    tensors[5i:5i+5] = (map_i, wg_base, bg_base, wb_base, bb_base) and
    cfg['maps'][i] = dict(K, wg_off, bg_off, wb_off, bb_off, nstride).
    cfg: dict(up, mode, training, eps, momentum, act, maps=[...]).
    """

    @staticmethod
    def forward(ctx, x, running_mean, running_var, cfg, *tensors):
        x = _c(x)
        tensors = [_c(t) for t in tensors]
        n, hs, ws, c = x.shape
        up, mode = cfg['up'], cfg['mode']
        h, w = hs * up, ws * up
        mean, rstd = _stats(x, n, hs * ws, c, mode, cfg['training'], running_mean, running_var, cfg['eps'],
                            cfg.get('momentum', 0.1), unbias_mul=up * up)
        d, arrays = SpadeFn._desc(cfg, n, h, w, c, tensors)
        out = torch.empty((n, h, w, c), device=x.device, dtype=torch.float32)
        fwd = lib.fsv_spade_fwd_tc if (CONV_USE_TC != 0 and lib.fsv_spade_fwd_tc_eligible(ctypes.byref(d))) else lib.fsv_spade_fwd
        _call(fwd, ctypes.byref(d), ptr(x), ptr(mean), ptr(rstd), *arrays, ptr(out), stream())
        ctx.cfg = cfg
        ctx.dims = (n, h, w, c)
        ctx.save_for_backward(x, mean, rstd, *tensors)
        return out

    @staticmethod
    def _desc(cfg, n, h, w, c, tensors):
        d = SpadeDesc()
        d.N, d.H, d.W, d.C, d.up, d.mode, d.act = n, h, w, c, cfg['up'], cfg['mode'], cfg.get('act', ACT_NONE)
        nm = len(cfg['maps'])
        d.nmaps = nm
        maps, wg, bg, wb, bb = PtrArray(), PtrArray(), PtrArray(), PtrArray(), PtrArray()
        for i, mc in enumerate(cfg['maps']):
            m, g0, g1, b0, b1 = tensors[5 * i:5 * i + 5]
            if tuple(m.shape[:3]) != (n, h, w):
                raise _lib.FsvError('SPADE label map %d has spatial size %s, expected %s (non-identity nearest resize is not '
                                    'supported)' % (i, tuple(m.shape[1:3]), (h, w)))
            d.K[i], d.m_ld[i], d.m_coff[i], d.w_nstride[i] = mc['K'], m.shape[3], 0, mc.get('nstride', 0)
            maps[i] = m.data_ptr()
            wg[i] = g0.data_ptr() + 4 * mc.get('wg_off', 0)
            bg[i] = None if g1 is None else g1.data_ptr() + 4 * mc.get('bg_off', 0)
            wb[i] = b0.data_ptr() + 4 * mc.get('wb_off', 0)
            bb[i] = None if b1 is None else b1.data_ptr() + 4 * mc.get('bb_off', 0)
        return d, (maps, wg, bg, wb, bb)

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        x, mean, rstd = saved[:3]
        tensors = list(saved[3:])
        cfg = ctx.cfg
        n, h, w, c = ctx.dims
        dout = _c(dout)
        dev = dout.device
        st = stream()
        nm = len(cfg['maps'])
        d, arrays = SpadeFn._desc(cfg, n, h, w, c, tensors)
        dxhat = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
        # fixed-weight maps get ONE interleaved (N,H,W,2C) [dgamma | dbeta] buffer so that their 1x1 data / weight gradients
        # are single GEMMs over 2C channels (tensor-core eligible); per-sample (hyper-weight) maps keep separate buffers
        # because their gamma / beta weights are not adjacent inside the hyper-network's flat output
        # plan per map: 'fixed' (shared weights), 'ps_tc' (per-sample hyper-weights, tcgen05 GEMMs), 'ps' (per-sample, SIMT)
        plan = []
        for i, mc in enumerate(cfg['maps']):
            if mc.get('nstride', 0) == 0:
                plan.append('fixed')
                continue
            K = mc['K']
            cdp = _conv_desc(n, h, w, K, 2 * c, 1, 1, 1, 0, 1, ACT_NONE, 1.0, 2 * c * K, 0, CONV_USE_TC)
            cdp.x_ld = tensors[5 * i].shape[3]
            ok = (CONV_USE_TC != 0 and lib.fsv_conv2d_dgrad_tc_eligible(ctypes.byref(cdp)) and
                  lib.fsv_conv2d_wgrad_tc_eligible(ctypes.byref(cdp)))
            plan.append('ps_tc' if ok else 'ps')
        fused = [pl != 'ps' for pl in plan]
        dgb, dgs, dbs = [None] * nm, [None] * nm, [None] * nm
        pg, pb = PtrArray(), PtrArray()
        for i in range(nm):
            if fused[i]:
                dgb[i] = torch.empty((n, h, w, 2 * c), device=dev, dtype=torch.float32)
                pg[i], pb[i] = dgb[i].data_ptr(), dgb[i].data_ptr() + 4 * c
                d.dgb_ld[i] = 2 * c
            else:
                dgs[i] = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
                dbs[i] = torch.empty((n, h, w, c), device=dev, dtype=torch.float32)
                pg[i], pb[i] = dgs[i].data_ptr(), dbs[i].data_ptr()
        bwd = lib.fsv_spade_bwd_tc if (CONV_USE_TC != 0 and lib.fsv_spade_fwd_tc_eligible(ctypes.byref(d))) else lib.fsv_spade_bwd
        _call(bwd, ctypes.byref(d), ptr(x), ptr(mean), ptr(rstd), *arrays, ptr(dout), ptr(dxhat), pg, pb, st)
        mode = cfg['mode']
        groups = n if mode == NORM_INSTANCE else 1
        batch_stats = 1 if (cfg['training'] or mode == NORM_INSTANCE) else 0
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            rpg = h * w if mode == NORM_INSTANCE else n * h * w
            scratch = torch.empty(2 * groups * c + int(lib.fsv_norm_work_doubles(groups, c, rpg)), device=dev, dtype=torch.float64)
            _call(lib.fsv_spade_norm_bwd, ptr(x), ptr(dxhat), ptr(mean), ptr(rstd), ptr(dx), ptr(scratch), n, h, w, c,
                  cfg['up'], mode, batch_stats, st)
        grads = [None] * len(tensors)
        gbuf = {}   # data_ptr -> gradient buffer, so that views of one flat hyper-output share one gradient

        def buf_for(idx):
            t = tensors[idx]
            key = (t.data_ptr(), t.numel())
            if key not in gbuf:
                gbuf[key] = torch.zeros_like(t)
                grads[idx] = gbuf[key]
            return gbuf[key]

        for i, mc in enumerate(cfg['maps']):
            m = tensors[5 * i]
            K = mc['K']
            need_m = ctx.needs_input_grad[4 + 5 * i]
            if plan[i] == 'ps_tc':
                # per-sample hyper-weights on tcgen05: gather this map's [Wgamma ; Wbeta] into one (B, 2C, K) tensor (tiny),
                # one data-gradient GEMM and one weight-gradient GEMM per map, then scatter dW back into the flat layout
                flat = tensors[5 * i + 1]
                go, bo = mc.get('wg_off', 0), mc.get('wb_off', 0)
                wcat = torch.cat([flat[:, go:go + c * K].reshape(n, c, K), flat[:, bo:bo + c * K].reshape(n, c, K)], 1).contiguous()
                cd = _conv_desc(n, h, w, K, 2 * c, 1, 1, 1, 0, 1, ACT_NONE, 1.0, 2 * c * K, 0, CONV_USE_TC)
                cd.x_ld = m.shape[3]
                if need_m:
                    dm = torch.empty_like(m)
                    _call(lib.fsv_conv2d_dgrad_tc, ctypes.byref(cd), ptr(dgb[i]), ptr(wcat.transpose(1, 2).contiguous()), ptr(dm), st)
                    grads[5 * i] = dm
                if ctx.needs_input_grad[4 + 5 * i + 1]:
                    dwcat = torch.empty((n, 2 * c, K), device=dev, dtype=torch.float32)
                    ws = torch.empty(int(lib.fsv_conv2d_wgrad_tc_workspace(ctypes.byref(cd))) // 4 + 64, device=dev, dtype=torch.float32)
                    _call(lib.fsv_conv2d_wgrad_tc, ctypes.byref(cd), ptr(m), ptr(dgb[i]), ptr(dwcat), ptr(ws), 0, st)
                    gflat = buf_for(5 * i + 1)
                    gflat[:, go:go + c * K] += dwcat[:, :c].reshape(n, c * K)
                    gflat[:, bo:bo + c * K] += dwcat[:, c:].reshape(n, c * K)
                continue
            if fused[i]:
                wg_t, bg_t, wb_t, bb_t = tensors[5 * i + 1:5 * i + 5]
                cd = _conv_desc(n, h, w, K, 2 * c, 1, 1, 1, 0, 1, ACT_NONE, 1.0, 0, 0, CONV_USE_TC)
                cd.x_ld = m.shape[3]
                wcat = torch.cat([wg_t.reshape(c, K), wb_t.reshape(c, K)], 0)            # (2C, K): tiny, parameter-side
                if need_m:
                    dm = torch.empty_like(m)
                    if cd.use_tc != 0 and lib.fsv_conv2d_dgrad_tc_eligible(ctypes.byref(cd)):
                        _call(lib.fsv_conv2d_dgrad_tc, ctypes.byref(cd), ptr(dgb[i]), ptr(wcat.t().contiguous()), ptr(dm), st)
                    else:
                        _call(lib.fsv_conv2d_dgrad, ctypes.byref(cd), ptr(dgb[i]), ptr(wcat), ptr(dm), 0, st)
                    grads[5 * i] = dm
                if ctx.needs_input_grad[4 + 5 * i + 1] or ctx.needs_input_grad[4 + 5 * i + 3]:
                    dwcat = torch.empty((2 * c, K), device=dev, dtype=torch.float32)
                    dbcat = torch.empty(2 * c, device=dev, dtype=torch.float32) if bg_t is not None else None
                    if cd.use_tc != 0 and lib.fsv_conv2d_wgrad_tc_eligible(ctypes.byref(cd)):
                        ws = torch.empty(int(lib.fsv_conv2d_wgrad_tc_workspace(ctypes.byref(cd))) // 4 + 1, device=dev, dtype=torch.float32)
                        _call(lib.fsv_conv2d_wgrad_tc, ctypes.byref(cd), ptr(m), ptr(dgb[i]), ptr(dwcat), ptr(ws), 0, st)
                        if dbcat is not None:
                            _call(lib.fsv_conv2d_wgrad, ctypes.byref(cd), ptr(m), ptr(dgb[i]), None, ptr(dbcat), 0, st)
                    else:
                        _call(lib.fsv_conv2d_wgrad, ctypes.byref(cd), ptr(m), ptr(dgb[i]), ptr(dwcat), ptr(dbcat), 0, st)
                    grads[5 * i + 1], grads[5 * i + 3] = dwcat[:c].reshape(wg_t.shape), dwcat[c:].reshape(wb_t.shape)
                    if dbcat is not None:
                        grads[5 * i + 2], grads[5 * i + 4] = dbcat[:c], dbcat[c:]
                continue
            ns = mc.get('nstride', 0)
            cd = _conv_desc(n, h, w, K, c, 1, 1, 1, 0, 1, ACT_NONE, 1.0, ns, ns, 0)
            cd.x_ld = m.shape[3]
            if need_m:
                dm = torch.empty_like(m) if m.shape[3] == K else torch.zeros_like(m)
                _call(lib.fsv_conv2d_dgrad, ctypes.byref(cd), ptr(dgs[i]), _off(tensors[5 * i + 1], mc.get('wg_off', 0)), ptr(dm), 0, st)
                _call(lib.fsv_conv2d_dgrad, ctypes.byref(cd), ptr(dbs[i]), _off(tensors[5 * i + 3], mc.get('wb_off', 0)), ptr(dm), 1, st)
                grads[5 * i] = dm
            if ctx.needs_input_grad[4 + 5 * i + 1]:
                gw = buf_for(5 * i + 1)
                gb = buf_for(5 * i + 2) if tensors[5 * i + 2] is not None else None
                _call(lib.fsv_conv2d_wgrad, ctypes.byref(cd), ptr(m), ptr(dgs[i]), _off(gw, mc.get('wg_off', 0)),
                      None if gb is None else _off(gb, mc.get('bg_off', 0)), 1, st)
            if ctx.needs_input_grad[4 + 5 * i + 3]:
                gw = buf_for(5 * i + 3)
                gb = buf_for(5 * i + 4) if tensors[5 * i + 4] is not None else None
                _call(lib.fsv_conv2d_wgrad, ctypes.byref(cd), ptr(m), ptr(dbs[i]), _off(gw, mc.get('wb_off', 0)),
                      None if gb is None else _off(gb, mc.get('bb_off', 0)), 1, st)
        return (dx, None, None, None) + tuple(grads)


# --------------------------------------------------------------------------- warp + composite

class WarpFn(torch.autograd.Function):
    """blend=False: out = [bilinear_warp(img, flow), mask] (N,H,W,Ci+1);  blend=True: raw*mask + warp*(1-mask)."""

    @staticmethod
    def forward(ctx, img, flow, mask, raw, blend):
        img, flow, mask, raw = _c(img), _c(flow), _c(mask), _c(raw)
        n, h, w, ci = img.shape
        co = ci if (blend or mask is None) else ci + 1
        out = torch.empty((n, h, w, co), device=img.device, dtype=torch.float32)
        _call(lib.fsv_warp_fwd, ptr(img), ptr(flow), ptr(mask), ptr(raw), ptr(out), n, h, w, ci, co, 0, 1 if blend else 0, stream())
        ctx.blend, ctx.co = blend, co
        ctx.save_for_backward(img, flow, mask, raw)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, flow, mask, raw = ctx.saved_tensors
        dout = _c(dout)
        n, h, w, ci = img.shape
        dflow = torch.empty_like(flow)
        dmask = torch.empty_like(mask) if (mask is not None and ctx.needs_input_grad[2]) else None
        draw = torch.empty_like(raw) if (raw is not None and ctx.needs_input_grad[3]) else None
        dimg = torch.zeros_like(img) if ctx.needs_input_grad[0] else None
        _call(lib.fsv_warp_bwd, ptr(img), ptr(flow), ptr(mask), ptr(raw), ptr(dout), ptr(dflow), ptr(dmask), ptr(draw), ptr(dimg),
              n, h, w, ci, ctx.co, 0, 1 if ctx.blend else 0, stream())
        return dimg, dflow, dmask, draw, None


def warp_concat(img, flow, mask):
    return WarpFn.apply(img, flow, mask, None, False)


def warp_blend(img, flow, mask, raw):
    return WarpFn.apply(img, flow, mask, raw, True)


# --------------------------------------------------------------------------- reference-feature outer product

class SoftmaxOuterFn(torch.autograd.Function):
    """out[b,c1,c2] = sum_hw img[b,hw,c1] * softmax_c(lab)[b,hw,c2]   (generator.py:381-388)."""

    @staticmethod
    def forward(ctx, img, lab):
        img, lab = _c(img), _c(lab)
        b, h, w, c = img.shape
        soft = torch.empty_like(lab)
        st = stream()
        _call(lib.fsv_softmax_rows_fwd, ptr(lab), ptr(soft), b * h * w, c, st)
        out = torch.empty((b, c, c), device=img.device, dtype=torch.float32)
        d = _conv_desc(b, h, w, c, c, 1, 1, 1, 0, 1, ACT_NONE, 1.0, c * c, 0, 0)
        _call(lib.fsv_conv2d_wgrad, ctypes.byref(d), ptr(soft), ptr(img), ptr(out), None, 0, st)
        ctx.save_for_backward(img, soft)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, soft = ctx.saved_tensors
        dout = _c(dout)
        b, h, w, c = img.shape
        st = stream()
        d = _conv_desc(b, h, w, c, c, 1, 1, 1, 0, 1, ACT_NONE, 1.0, c * c, 0, 0)
        dimg = dlab = None
        if ctx.needs_input_grad[0]:
            dimg = torch.empty_like(img)
            _call(lib.fsv_conv2d_fwd, ctypes.byref(d), ptr(soft), ptr(dout), None, None, ptr(dimg), st)
        if ctx.needs_input_grad[1]:
            dsoft = torch.empty_like(soft)
            _call(lib.fsv_conv2d_dgrad, ctypes.byref(d), ptr(img), ptr(dout), ptr(dsoft), 0, st)
            dlab = torch.empty_like(soft)
            _call(lib.fsv_softmax_rows_bwd, ptr(soft), ptr(dsoft), ptr(dlab), b * h * w, c, st)
        return dimg, dlab


def softmax_outer(img_feat, lab_feat):
    return SoftmaxOuterFn.apply(img_feat, lab_feat)


# --------------------------------------------------------------------------- K-shot attention helpers (generator.py:298-316)

class SoftmaxChannelsFn(torch.autograd.Function):
    """softmax over the channel (last NHWC) axis: fsv_softmax_rows_fwd/bwd."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        _call(lib.fsv_softmax_rows_fwd, ptr(x), ptr(y), x.numel() // x.shape[-1], x.shape[-1], stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(y)
        _call(lib.fsv_softmax_rows_bwd, ptr(y), ptr(dy), ptr(dx), y.numel() // y.shape[-1], y.shape[-1], stream())
        return dx


def softmax_channels(x):
    return SoftmaxChannelsFn.apply(x)


class PerSampleMatmulFn(torch.autograd.Function):
    """y[b, h, w, :] = W_b @ x[b, h, w, :], W_b = flat[b].view(cout, cin): the two GEMMs of the K-shot attention (generator.py:298-316).
    On the tcgen05 path the forward IS the per-sample data-gradient kernel with the roles of the channel axes swapped
    (dx[ci'] = sum_co' dy[co'] wt[ci'][co'] with ci' = cout, co' = cin and wt = W as stored), the backward uses the per-sample
    data-gradient / weight-gradient kernels directly; shapes they do not serve (tiny images, odd widths) and the exact-fp32 mode
    (CONV_USE_TC = 0) run the SIMT kernels."""

    @staticmethod
    def forward(ctx, x, flat, cout, cin):
        x, flat = _c(x), _c(flat)
        b, h, w, _ = x.shape
        st = stream()
        y = torch.empty((b, h, w, cout), device=x.device, dtype=torch.float32)
        dfw = _conv_desc(b, h, w, cout, cin, 1, 1, 1, 0, 1, ACT_NONE, 1.0, cout * cin, 0, CONV_USE_TC)     # swapped roles (see above)
        if CONV_USE_TC != 0 and lib.fsv_conv2d_dgrad_tc_eligible(ctypes.byref(dfw)):
            _call(lib.fsv_conv2d_dgrad_tc, ctypes.byref(dfw), ptr(x), ptr(flat), ptr(y), st)
        else:
            d = _conv_desc(b, h, w, cin, cout, 1, 1, 1, 0, 1, ACT_NONE, 1.0, cout * cin, 0, 0)
            _call(lib.fsv_conv2d_fwd, ctypes.byref(d), ptr(x), ptr(flat), None, None, ptr(y), st)
        ctx.save_for_backward(x, flat)
        ctx.dims = (b, h, w, cout, cin)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, flat = ctx.saved_tensors
        b, h, w, cout, cin = ctx.dims
        dy = _c(dy)
        st = stream()
        d = _conv_desc(b, h, w, cin, cout, 1, 1, 1, 0, 1, ACT_NONE, 1.0, cout * cin, 0, CONV_USE_TC)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if CONV_USE_TC != 0 and lib.fsv_conv2d_dgrad_tc_eligible(ctypes.byref(d)):
                wt = flat.view(b, cout, cin).transpose(1, 2).contiguous()
                _call(lib.fsv_conv2d_dgrad_tc, ctypes.byref(d), ptr(dy), ptr(wt), ptr(dx), st)
            else:
                d0 = _conv_desc(b, h, w, cin, cout, 1, 1, 1, 0, 1, ACT_NONE, 1.0, cout * cin, 0, 0)
                _call(lib.fsv_conv2d_dgrad, ctypes.byref(d0), ptr(dy), ptr(flat), ptr(dx), 0, st)
        if ctx.needs_input_grad[1]:
            if CONV_USE_TC != 0 and lib.fsv_conv2d_wgrad_tc_eligible(ctypes.byref(d)):
                dw = torch.empty_like(flat)
                ws = torch.empty(int(lib.fsv_conv2d_wgrad_tc_workspace(ctypes.byref(d))) // 4 + 64, device=dy.device, dtype=torch.float32)
                _call(lib.fsv_conv2d_wgrad_tc, ctypes.byref(d), ptr(x), ptr(dy), ptr(dw), ptr(ws), 0, st)
            else:
                dw = torch.zeros_like(flat)
                d0 = _conv_desc(b, h, w, cin, cout, 1, 1, 1, 0, 1, ACT_NONE, 1.0, cout * cin, 0, 0)
                _call(lib.fsv_conv2d_wgrad, ctypes.byref(d0), ptr(x), ptr(dy), ptr(dw), None, 1, st)
        return dx, dw, None, None


def per_sample_matmul(x, flat, cout, cin):
    """y[b, h, w, :] = W_b @ x[b, h, w, :] with W_b = flat[b].view(cout, cin): a per-sample 1x1 conv without bias (the two
    GEMMs of the attention module: key^T query and x_ref @ attention).  Gradients flow to both x and flat."""
    assert x.shape[3] == cin and flat.shape[1] == cout * cin
    return PerSampleMatmulFn.apply(x, flat, cout, cin)


# --------------------------------------------------------------------------- pose label preprocessing + face region
def _plane(label, ch):
    """device address and batch stride of channel ``ch`` of a contiguous NCHW label."""
    label = _c(label)
    b, c, h, w = label.shape
    return label, c_vp(label.data_ptr() + 4 * (ch % c) * h * w), c * h * w


def fg_mask(label, ch=2, thr=-1.0):
    """get_fg_mask (models/input_process.py:52-61): (MaxPool2d(15,1,7)(label[:, ch]) > thr).float() -> (B, 1, H, W)."""
    label, p, ns = _plane(label, ch)
    b, _, h, w = label.shape
    out = torch.empty((b, 1, h, w), device=label.device, dtype=torch.float32)
    _call(lib.fsv_fg_mask, p, ns, ptr(out), b, h, w, float(thr), stream())
    return out


def face_mask_avg15(label, ch=2):
    """AvgPool2d(15,1,7)(get_face_mask(label[:, ch])) (input_process.py:81-93, loss_collector.py:178-179) -> (B, 1, H, W)."""
    label, p, ns = _plane(label, ch)
    b, _, h, w = label.shape
    out = torch.empty((b, 1, h, w), device=label.device, dtype=torch.float32)
    _call(lib.fsv_face_mask_avg15, p, ns, ptr(out), b, h, w, stream())
    return out


def part_masks(label, ch=2):
    """get_part_mask (input_process.py:63-79) of label[:, ch] -> NHWC (B, H, W, 9)."""
    label, p, ns = _plane(label, ch)
    b, _, h, w = label.shape
    out = torch.empty((b, h, w, 9), device=label.device, dtype=torch.float32)
    _call(lib.fsv_part_masks, p, ns, ptr(out), b, h, w, stream())
    return out


def face_bbox(planes, thr, openpose, crop_smaller=0):
    """get_face_region (models/face_refiner.py:52-83) on the device: ``planes`` = up to three (B, 1|*, H, W)-addressable
    (tensor, channel) pairs that must all exceed ``thr``.  -> int32 (B, 4) = ys, ye, xs, xe, never copied to the host."""
    ps = [_plane(t, ch) for t, ch in planes]
    b, _, h, w = ps[0][0].shape
    box = torch.empty((b, 4), device=ps[0][0].device, dtype=torch.int32)
    a = ps + [(None, None, 0)] * (3 - len(ps))
    _call(lib.fsv_face_bbox, a[0][1], a[1][1], a[2][1], a[0][2], a[1][2], a[2][2], float(thr), b, h, w, 1 if openpose else 0,
          int(crop_smaller), c_vp(box.data_ptr()), stream())
    return box


class CropResizeFn(torch.autograd.Function):
    """crop_face_region (face_refiner.py:34-38): image (B, C<=4, H, W) NCHW-shaped (any strides: NCHW tensors and NCHW views
    of NHWC buffers alike) -> NHWC (B, S, S, C), boxes (B, 4) int32 on the device."""

    @staticmethod
    def forward(ctx, image, box, size):
        _lib.require_cuda(image, box)
        if image.dtype != torch.float32 or box.dtype != torch.int32:
            raise _lib.FsvError('crop_resize: float32 image and int32 boxes expected')
        b, c, h, w = image.shape
        sn, sc, sh, sw = image.stride()
        out = torch.empty((b, size, size, c), device=image.device, dtype=torch.float32)
        _call(lib.fsv_crop_resize_fwd, ptr(image), sn, sc, sh, sw, c_vp(box.data_ptr()), ptr(out), b, c, size, c, 0, stream())
        ctx.save_for_backward(box)
        ctx.dims = (b, c, h, w, size)
        return out

    @staticmethod
    def backward(ctx, dout):
        (box,) = ctx.saved_tensors
        b, c, h, w, size = ctx.dims
        dout = _c(dout)
        dimg = torch.empty((b, h, w, c), device=dout.device, dtype=torch.float32)
        _call(lib.fsv_crop_resize_bwd, ptr(dout), c, 0, c_vp(box.data_ptr()), ptr(dimg), b, c, h, w, size, stream())
        return dimg.permute(0, 3, 1, 2), None, None


def crop_resize(image, box, size):
    """-> NHWC (B, size, size, C)."""
    return CropResizeFn.apply(image, box, size)


# --------------------------------------------------------------------------- discriminator input packing
def pad_channels(c):
    """Channel count a network-input buffer is padded to: thin-layer kernels serve Cin <= 8 as is; wider inputs (pose flow net 15,
    pose D 20, street label 20 / D 46) are zero-padded to a multiple of 32 so that their first conv runs on the tcgen05 path (its
    K block is 32 channels); the consumer conv zero-pads its weight to match (layers.Conv2d), so values are unchanged."""
    return c if c <= 8 else (c + 31) // 32 * 32


def pack_rows(dst, row0, tensors, coff):
    """NCHW tensors -> consecutive channel slices (from ``coff``) of rows [row0, row0 + B) of the NHWC buffer ``dst``.
    Returns the next free channel offset.  No autograd (constant inputs)."""
    for t in tensors:
        t = _c(t.detach())
        b, c, h, w = t.shape
        _call(lib.fsv_nchw_to_nhwc, ptr(t), c_vp(dst.data_ptr() + 4 * row0 * h * w * dst.shape[3]), b, c, h, w, dst.shape[3], coff, stream())
        coff += c
    return coff


class DInputFn(torch.autograd.Function):
    """The discriminator input of loss_collector.py:47-58: rows [fake ; real], channels [ref.. | label.. | image] in ONE NHWC
    buffer.  Everything but the fake image is constant over an iteration and was packed once into ``base`` (2B, H, W, Cp);
    a call copies ``base`` and drops the fake frame (an NCHW-shaped view of the generator's NHWC output) into rows [0, B) at
    channel ``coff``.  Gradient: that channel slice of rows [0, B)."""

    @staticmethod
    def forward(ctx, base, fake, coff):
        x = base.clone()
        b, c, h, w = fake.shape
        f = _c(fake.permute(0, 2, 3, 1))
        _call(lib.fsv_copy_channels, ptr(f), c, 0, ptr(x), x.shape[3], coff, b * h * w, c, 0, stream())
        ctx.dims = (b, c, h, w, coff)
        return x

    @staticmethod
    def backward(ctx, dx):
        b, c, h, w, coff = ctx.dims
        dx = _c(dx)
        g = torch.empty((b, h, w, c), device=dx.device, dtype=torch.float32)
        _call(lib.fsv_copy_channels, ptr(dx), dx.shape[3], coff, ptr(g), c, 0, b * h * w, c, 0, stream())
        return None, g.permute(0, 3, 1, 2), None


def d_input(base, fake, coff):
    return DInputFn.apply(base, fake, coff)


# --------------------------------------------------------------------------- fused generator-step losses
class FlowMaskLossFn(torch.autograd.Function):
    """loss_collector.py:131-204 in one forward and one backward pass over the frame (csrc/losses.cu).  Inputs (None = absent):
    warp0, mask0, warp1, mask1 (NHWC), tgt (NCHW), fake (NHWC), ref_body_warp, body (NHWC, 9 ch), ref_fg_warp, fg, face_avg, fg_diff
    (one channel).  -> tensor (2,) = [F_Warp / lambda_flow, F_Mask / lambda_mask]."""

    NAMES = ('warp0', 'mask0', 'warp1', 'mask1', 'tgt', 'fake', 'ref_body_warp', 'body', 'ref_fg_warp', 'fg', 'face_avg', 'fg_diff')

    @staticmethod
    def forward(ctx, *ts):
        ts = [_c(t.detach()) if t is not None else None for t in ts]
        tgt = ts[4]
        b, _, h, w = tgt.shape
        d = _lib.FlowMaskDesc()
        for name, t in zip(FlowMaskLossFn.NAMES, ts):
            setattr(d, name, None if t is None else t.data_ptr())
        d.B, d.H, d.W = b, h, w
        out = torch.empty(2, device=tgt.device, dtype=torch.float32)
        work = torch.empty(int(lib.fsv_flow_mask_loss_work_doubles()), device=tgt.device, dtype=torch.float64)
        _call(lib.fsv_flow_mask_loss_fwd, ctypes.byref(d), ptr(out), ptr(work), stream())
        ctx.save_for_backward(*[t for t in ts if t is not None])
        ctx.present = [t is not None for t in ts]
        ctx.dims = (b, h, w)
        return out

    @staticmethod
    def backward(ctx, gout):
        it = iter(ctx.saved_tensors)
        ts = [next(it) if pr else None for pr in ctx.present]
        b, h, w = ctx.dims
        d = _lib.FlowMaskDesc()
        for name, t in zip(FlowMaskLossFn.NAMES, ts):
            setattr(d, name, None if t is None else t.data_ptr())
        d.B, d.H, d.W = b, h, w
        gout = _c(gout)
        dev = gout.device

        def buf(t, want=True):
            return torch.empty_like(t) if (t is not None and want) else None
        dw0, dm0, dw1, dm1 = buf(ts[0]), buf(ts[1]), buf(ts[2]), buf(ts[3])
        dfake = buf(ts[5], ts[10] is not None)
        drbw, drfw = buf(ts[6]), buf(ts[8])
        _call(lib.fsv_flow_mask_loss_bwd, ctypes.byref(d), ptr(gout), _off(gout, 1), ptr(dw0), ptr(dw1), ptr(dm0), ptr(dm1), ptr(dfake), ptr(drbw),
              ptr(drfw), stream())
        return dw0, dm0, dw1, dm1, None, dfake, drbw, None, drfw, None, None, None


def flow_mask_losses(warp0, mask0, warp1, mask1, tgt, fake, ref_body_warp, body, ref_fg_warp, fg, face_avg, fg_diff):
    return FlowMaskLossFn.apply(warp0, mask0, warp1, mask1, tgt, fake, ref_body_warp, body, ref_fg_warp, fg, face_avg, fg_diff)


class HalvesL1Fn(torch.autograd.Function):
    """mean |x[:B] - x[B:].detach()| for a contiguous tensor whose first dimension is the batch [fake ; real] (loss_collector.py:206-215)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        half = x.numel() // 2
        out = torch.empty(1, device=x.device, dtype=torch.float32)
        work = torch.empty(1024, device=x.device, dtype=torch.float64)
        _call(lib.fsv_halves_l1_fwd, ptr(x), half, ptr(out), ptr(work), stream())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _c(g)
        dx = torch.empty_like(x)
        _call(lib.fsv_halves_l1_bwd, ptr(x), x.numel() // 2, ptr(g), ptr(dx), stream())
        return dx


def halves_l1(x):
    return HalvesL1Fn.apply(x)
