"""Parameter-holding layers of the B200 hot path.

Every class keeps the attribute names (hence ``state_dict`` keys) of the reference module it
stands in for, so reference checkpoints load unchanged; the forward bodies
call the fsv C ABI through ``fsv.ops`` on NHWC activations.  Spectral normalisation keeps the
``torch.nn.utils.spectral_norm`` registration (parameter ``weight_orig``, buffers ``weight_u/_v``, state_dict
hooks) exactly as the reference applies it, but the per-forward weight computation (one power iteration in
training mode) runs in ``fsv_spectral_fwd/bwd`` instead of the ~20-launch torch hook.
"""
import os

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm as _sn
from torch.nn.utils.spectral_norm import SpectralNorm as _SNHook

from .. import ops
from ..ops import ACT_NONE, ACT_LRELU, NORM_BATCH, NORM_INSTANCE


# one grouped spectral-norm launch per network forward instead of one per module (see SpectralPlanner); False = per-module calls
GROUP_SPECTRAL = os.environ.get('FSV_GROUP_SPECTRAL', '1') != '0'
# Number of groups a large network's spectral weights are split into (contiguous chunks of the call order, balanced by size).  With
# several chunks the grouped backward of a chunk runs as soon as that chunk's layers are done, so its weight gradients reach the
# data-parallel all-reduce buckets earlier; measured at N=2 that bought nothing over one group (the buckets
# fired on the side stream already overlap), so the default stays 1 (FSV_SPECTRAL_CHUNKS overrides).
SPECTRAL_GROUP_CHUNKS = int(os.environ.get('FSV_SPECTRAL_CHUNKS', '1'))
SPECTRAL_CHUNK_MIN_NUMEL = int(os.environ.get('FSV_SPECTRAL_CHUNK_MIN', '20000000'))      # smaller networks (the discriminators) stay in one group


class SpectralPlanner:
    """Per top-level network (generator, discriminator): computes the spectral weights of ALL modules a forward pass will call in
    three kernel launches up front (ops.SpectralGroup) instead of three launches per module.

    Which modules a forward calls depends on its arguments (temporal inputs, eval-time weight cache, K-shot attention), and
    torch.nn.utils.spectral_norm semantics must be kept exactly: ``weight_u/_v`` advance once per module CALL in training mode,
    and only for modules that are called.  So the first forward with a given signature runs the ordinary per-module path and
    records the sequence of spectral modules it touched; later forwards with the same signature compute that recorded set in one
    group at entry, and each module picks its weight up when it runs.  A module called a second time in the same forward (the
    shared flow network of the temporal phase, generator.py:159,166) falls back to its own power iteration, as does any module
    the recording missed."""

    def __init__(self, net):
        self.plans = {}          # signature -> dict(mods=[(module, want_wt)], group=SpectralGroup or None)
        self.ready = None
        self.recording = None
        self.active = False
        for m in net.modules():
            if getattr(m, 'fsv_spectral', False):
                m._sn_planner = self

    def begin(self, sig):
        if not GROUP_SPECTRAL or self.active:
            return False
        self.active = True
        plan = self.plans.get(sig)
        self.sig = sig
        if plan is None:
            self.recording, self.ready = [], {}
            return True
        mods = plan['mods']
        self.recording = None
        if not mods:
            self.ready = {}
            return True
        if plan.get('chunks') is None:
            total = sum(m.weight_orig.numel() for m, _ in mods)
            k = SPECTRAL_GROUP_CHUNKS if total > SPECTRAL_CHUNK_MIN_NUMEL else 1          # only the generator is worth splitting
            chunks, cur, acc = [], [], 0
            for m, want in mods:
                cur.append((m, want))
                acc += m.weight_orig.numel()
                if len(chunks) < k - 1 and acc >= total * (len(chunks) + 1) / k:
                    chunks.append(cur)
                    cur = []
            if cur:
                chunks.append(cur)
            plan['chunks'], plan['groups'] = chunks, [None] * len(chunks)
        self.ready = {}
        for ci, chunk in enumerate(plan['chunks']):
            entries = [(m.weight_orig, m.weight_u, m.weight_v, want) for m, want in chunk]
            if plan['groups'][ci] is None or not plan['groups'][ci].matches(entries):
                plan['groups'][ci] = ops.SpectralGroup(entries)
            m0 = chunk[0][0]
            ws, wts = ops.spectral_group_weights(plan['groups'][ci], m0.training, m0.fsv_spectral_eps, [e[0] for e in entries])
            self.ready.update({id(m): (w, wt) for (m, _), w, wt in zip(chunk, ws, wts)})
        return True

    def end(self):
        if self.recording is not None:
            seen, mods = set(), []
            for m, want in self.recording:
                if id(m) not in seen:
                    seen.add(id(m))
                    mods.append((m, want))
            self.plans[self.sig] = dict(mods=mods, group=None)
        self.recording, self.ready, self.active = None, None, False

    def get(self, module, want_wt):
        """-> (w_sn OHWI, wt or None)"""
        if self.active and self.ready:
            hit = self.ready.pop(id(module), None)
            if hit is not None:
                return hit          # (a missing channel-swapped copy is rebuilt by the conv's backward itself)
        if self.active and self.recording is not None:
            self.recording.append((module, bool(want_wt)))
        return _spectral_single(module, want_wt)


def _spectral_single(module, want_wt):
    if want_wt:
        return ops.spectral_weight(module.weight_orig, module.weight_u, module.weight_v, module.training, module.fsv_spectral_eps, want_wt=True)
    return ops.spectral_weight(module.weight_orig, module.weight_u, module.weight_v, module.training, module.fsv_spectral_eps), None


def spectral_weight_of(module, want_wt):
    planner = getattr(module, '_sn_planner', None)
    if planner is not None:
        return planner.get(module, want_wt)
    return _spectral_single(module, want_wt)


class Conv2d(nn.Module):
    """nn.Conv2d stand-in (weight (Cout,Cin,kh,kw), optional bias) running fsv conv kernels."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty(cout, cin, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def ohwi(self, want_wt=False):
        """OHWI weight [and its channel-swapped copy for the tcgen05 data gradient, or None]."""
        if getattr(self, 'fsv_spectral', False):
            w, wt = spectral_weight_of(self, want_wt)
            return (w, wt) if want_wt else w
        w = ops.to_ohwi(self.weight)
        return (w, None) if want_wt else w

    def forward(self, x, up=1, act=ACT_NONE, residual=None, out_scale=1.0, in_act=ACT_NONE):
        cin = x.shape[3]
        # the swapped copy only pays off where the tensor-core data gradient will run (conv_tc.cu eligibility)
        want = cin % 16 == 0 and self.out_channels % 32 == 0 and x.requires_grad
        w, wt = self.ohwi(want_wt=True) if want else (self.ohwi(), None)
        if cin != self.in_channels:
            # network-input buffers are zero-padded to the tcgen05 K block (ops.pad_channels): pad the weight's input-channel axis
            # with zeros to match -- same values; autograd slices the padded weight gradient back
            if cin < self.in_channels:
                raise ValueError('conv input has %d channels, expected %d' % (cin, self.in_channels))
            w = torch.nn.functional.pad(w, (0, cin - self.in_channels))
            if wt is not None:
                wt = torch.nn.functional.pad(wt, (0, 0, 0, 0, 0, 0, 0, cin - self.in_channels))
        if not self.__dict__.get('_guarded'):
            for p in self.parameters(recurse=False):
                ops.guard_param(p)
            self.__dict__['_guarded'] = True
        # weight / bias gradients of this conv are consumed by the spectral backward or OhwiFn (both continue on the side stream) and
        # then adopted by the leaves: safe to compute off the critical path.  Not so for a zero-padded weight (its gradient is sliced
        # by torch on the main stream).
        return ops.conv2d(x, w, self.bias, stride=self.stride, pad=self.padding, up=up, act=act,
                          out_scale=out_scale, residual=residual, in_act=in_act, wt=wt, side_ok=(cin == self.in_channels))


class Linear(nn.Module):
    """nn.Linear stand-in for the hyper-network MLPs (generator.py:103-110)."""

    def __init__(self, fin, fout):
        super().__init__()
        self.in_features, self.out_features = fin, fout
        self.weight = nn.Parameter(torch.empty(fout, fin))
        self.bias = nn.Parameter(torch.zeros(fout))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x, act=ACT_NONE):
        wt = None
        if getattr(self, 'fsv_spectral', False):
            want = self.in_features % 16 == 0 and self.out_features % 32 == 0 and x.requires_grad
            w, wt = spectral_weight_of(self, want)
        else:
            w = self.weight
        if not self.__dict__.get('_guarded'):
            for p in self.parameters(recurse=False):
                ops.guard_param(p)
            self.__dict__['_guarded'] = True
        return ops.linear(x, w, self.bias, act=act, wt=wt, side_ok=bool(getattr(self, 'fsv_spectral', False)))


def spectral(module):
    """torch.nn.utils.spectral_norm(module) as the reference calls it (default name 'weight', 1 power iteration, eps
    1e-12, dim 0): same parameter / buffer registration, initial u / v draw and state_dict hooks; the forward-pre-hook
    that recomputes ``module.weight`` is dropped -- Conv2d / Linear call ``ops.spectral_weight`` on
    ``weight_orig/_u/_v`` themselves (``module.weight`` is therefore a stale construction-time tensor: do not read it)."""
    _sn(module)
    for key, hook in list(module._forward_pre_hooks.items()):
        if isinstance(hook, _SNHook):
            module.fsv_spectral_eps = hook.eps
            del module._forward_pre_hooks[key]
    module.fsv_spectral = True
    return module


class BatchNorm(nn.Module):
    """(Sync)BatchNorm2d stand-in: local batch statistics, eps 1e-5, momentum 0.1."""

    def __init__(self, c, affine=True, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum, self.affine = c, eps, momentum, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(c))
            self.bias = nn.Parameter(torch.zeros(c))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def tick(self):
        if self.training:
            self.num_batches_tracked += 1

    def forward(self, x, act=ACT_NONE):
        self.tick()
        return ops.norm_act(x, self.weight, self.bias, self.running_mean, self.running_var, NORM_BATCH, self.training,
                            self.eps, self.momentum, act)


class InstanceNorm(nn.Module):
    """nn.InstanceNorm2d(affine, eps=0.1) stand-in (normalization.py:35,82); no running statistics."""

    def __init__(self, c, affine=True, eps=0.1):
        super().__init__()
        self.num_features, self.eps, self.affine = c, eps, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(c))
            self.bias = nn.Parameter(torch.zeros(c))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x, act=ACT_NONE):
        return ops.norm_act(x, self.weight, self.bias, None, None, NORM_INSTANCE, self.training, self.eps, 0.1, act)


class SPADE(nn.Module):
    """normalization.py:18-52.  ``norm`` holds the BatchNorm buffers (or nothing for instance norm),
    ``mlp_gamma{s}/mlp_beta{s}`` the fixed 1x1 weights (absent for map 0 when ``params_free``)."""

    def __init__(self, norm_nc, hidden_nc, norm='batch', ks=1, params_free=False):
        super().__init__()
        if ks != 1:
            raise NotImplementedError('SPADE with spade_ks != 1 is outside the hot-path scope (default spade_ks=1)')
        if not isinstance(hidden_nc, list):
            hidden_nc = [hidden_nc]
        self.norm_nc, self.hidden_nc, self.params_free = norm_nc, list(hidden_nc), params_free
        for i, nh in enumerate(hidden_nc):
            if not params_free or i != 0:
                s = str(i + 1) if i > 0 else ''
                setattr(self, 'mlp_gamma%s' % s, Conv2d(nh, norm_nc, 1))
                setattr(self, 'mlp_beta%s' % s, Conv2d(nh, norm_nc, 1))
        self.batch = 'batch' in norm
        self.norm = BatchNorm(norm_nc, affine=False) if self.batch else InstanceNorm(norm_nc, affine=False)

    def forward(self, x, maps, weights=None, up=1, act=ACT_NONE):
        """x: (N, H/up, W/up, C) NHWC; maps: list of NHWC label maps at (H, W) or None;
        weights: None or (flat, wg_off, bg_off, wb_off, bb_off) locating the hyper-weights of map 0."""
        if not isinstance(maps, (list, tuple)):
            maps = [maps]
        C = self.norm_nc
        tensors, mcfg = [], []
        for i, m in enumerate(maps):
            if m is None:
                continue
            K = m.shape[3]
            if weights is None or i != 0:
                s = str(i + 1) if i > 0 else ''
                g, b = getattr(self, 'mlp_gamma%s' % s), getattr(self, 'mlp_beta%s' % s)
                tensors += [m, g.weight.reshape(C, K), g.bias, b.weight.reshape(C, K), b.bias]
                mcfg.append(dict(K=K))
            else:
                # normalization.py:48-50 calls batch_conv(m, weights[0][j]) with j = min(i, 1) = 0, i.e. with the
                # weight tensor of the [weight, bias] pair only: the generated bias slots are never applied.
                flat, wg_off, _bg_off, wb_off, _bb_off = weights
                tensors += [m, flat, None, flat, None]
                mcfg.append(dict(K=K, wg_off=wg_off, wb_off=wb_off, nstride=flat.shape[1]))
        if self.batch:
            self.norm.tick()
        cfg = dict(up=up, mode=NORM_BATCH if self.batch else NORM_INSTANCE, training=self.training,
                   eps=self.norm.eps, momentum=0.1, act=act, maps=mcfg)
        rm = self.norm.running_mean if self.batch else None
        rv = self.norm.running_var if self.batch else None
        return ops.SpadeFn.apply(x, rm, rv, cfg, *tensors)


class ConvNormAct(nn.Module):
    """architecture.py:57-69 SPADEConv2d with a plain norm: ``conv`` (spectral, bias) -> ``bn`` -> LeakyReLU."""

    def __init__(self, fin, fout, stride=1):
        super().__init__()
        self.conv = spectral(Conv2d(fin, fout, 3, stride=stride, padding=1))
        self.bn = BatchNorm(fout, affine=True)

    def forward(self, x):
        return self.bn(self.conv(x), act=ACT_LRELU)


class SPADEResnetBlock(nn.Module):
    """architecture.py:71-108.  With a 'spade' norm this is the generator main-branch block; with a plain
    batch norm (flow network, generator.py:477-480) ``bn_*`` are BatchNorm layers."""

    def __init__(self, fin, fout, norm='batch', hidden_nc=0, norm_params_free=False):
        super().__init__()
        fhidden = min(fin, fout)
        self.learned_shortcut = fin != fout
        self.spade = 'spade' in norm
        self.conv_0 = spectral(Conv2d(fin, fhidden, 3, padding=1))
        self.conv_1 = spectral(Conv2d(fhidden, fout, 3, padding=1))
        if self.learned_shortcut:
            self.conv_s = spectral(Conv2d(fin, fout, 1, bias=False))
        if self.spade:
            mk = lambda c: SPADE(c, hidden_nc, norm=norm, ks=1, params_free=norm_params_free)  # noqa: E731
        else:
            mk = lambda c: BatchNorm(c, affine=True)  # noqa: E731
        self.bn_0 = mk(fin)
        self.bn_1 = mk(fhidden)
        if self.learned_shortcut:
            self.bn_s = mk(fin)

    def forward(self, x, maps=None, norm_weights=None, up=1):
        """x: NHWC at 1/up of the block's resolution (the reference's nearest upsample, generator.py:207,
        is folded into the SPADE loads)."""
        if not self.spade:
            dx = self.conv_0(self.bn_0(x, act=ACT_LRELU))
            return self.conv_1(self.bn_1(dx, act=ACT_LRELU), residual=x)
        nw = norm_weights if norm_weights is not None else [None] * 3
        if self.learned_shortcut:
            xs = self.conv_s(self.bn_s(x, maps, nw[2], up=up, act=ACT_NONE))
        else:
            xs = x if up == 1 else ops.upsample2x(x)
        dx = self.conv_0(self.bn_0(x, maps, nw[0], up=up, act=ACT_LRELU))
        return self.conv_1(self.bn_1(dx, maps, nw[1], up=1, act=ACT_LRELU), residual=xs)


def init_weights(net, init_type='xavier', gain=0.02, reach_spectral=True):
    """base_network.py:86-115: the requested init (default xavier-normal, gain 0.02) on every conv / linear weight, zero biases;
    norm-layer affine parameters stay at (1, 0) -- the reference's 'BatchNorm2d' name test does not match its SyncBatchNorm /
    InstanceNorm classes.

    ``reach_spectral``: the reference writes through ``m.weight.data``, which aliases ``weight_orig`` of a spectral-normalised
    module only as long as the module has not been moved: ``define_G`` / ``define_D`` call ``.cuda()`` BEFORE ``init_weights``
    (networks/__init__.py:35-38,51-54), after which ``m.weight`` is a detached copy and the init never reaches ``weight_orig`` --
    on a GPU the reference's spectral layers therefore keep PyTorch's default kaiming-uniform init (their biases are still
    zeroed).  The forward is invariant to that scale (W / sigma) but Adam's relative step on ``weight_orig`` is not, so the
    factories mirror the effective behaviour: ``reach_spectral = (no GPU move happened)``.  ``init_temporal_network`` builds and
    initialises its new sub-networks before they are moved (generator.py:160-172), there the init does reach ``weight_orig``."""
    for m in net.modules():
        if isinstance(m, (Conv2d, Linear)):
            spectral_m = hasattr(m, 'weight_orig')
            w = m.weight_orig if spectral_m else m.weight
            if spectral_m and not reach_spectral:
                pass
            elif init_type == 'normal':
                nn.init.normal_(w.data, 0.0, gain)
            elif init_type == 'xavier':
                nn.init.xavier_normal_(w.data, gain=gain)
            elif init_type == 'xavier_uniform':
                nn.init.xavier_uniform_(w.data, gain=1.0)
            elif init_type == 'kaiming':
                nn.init.kaiming_normal_(w.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                nn.init.orthogonal_(w.data, gain=gain)
            elif init_type == 'none':
                pass
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)
