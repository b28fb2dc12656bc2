"""MultiscaleDiscriminator / NLayerDiscriminator on the fsv_b200 kernels.

Drop-in for models/networks/discriminator.py:16-102 of the reference (``netD_subarch='n_layers'``,
norm 'spectralinstance'): same constructor arguments, ``forward(input, ref=None)`` returning
list[num_D] of list[n_layers+2] feature tensors (NCHW-shaped), same state_dict names.
"""
import numpy as np
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_LRELU
from .layers import Conv2d, InstanceNorm, BatchNorm, SpectralPlanner, spectral
from .generator import BaseNetwork, _Act


def get_nonspade_norm_layer(opt, norm_type='instance'):
    """normalization.py:54-88: returns a function wrapping a conv into Sequential(spectral(conv) w/o bias, norm)."""
    def add_norm_layer(layer):
        sub = norm_type
        if norm_type.startswith('spectral'):
            layer = spectral(layer)
            sub = norm_type[len('spectral'):]
        if sub == 'none' or len(sub) == 0:
            return layer
        if getattr(layer, 'bias', None) is not None:
            delattr(layer, 'bias')
            layer.register_parameter('bias', None)
        if sub in ('batch', 'syncbatch'):
            norm = BatchNorm(layer.out_channels, affine=True)
        elif sub == 'instance':
            norm = InstanceNorm(layer.out_channels, affine=True, eps=0.1)
        else:
            raise ValueError('normalization layer %s is not recognized' % sub)
        return nn.Sequential(layer, norm)
    return add_norm_layer


class _planned:
    """One grouped spectral-norm launch for the whole discriminator forward (layers.SpectralPlanner)."""

    def __init__(self, net, x):
        import torch
        self.p = net.__dict__.get('_planner')
        if self.p is None:
            self.p = net.__dict__['_planner'] = SpectralPlanner(net)
        self.sig = (net.training, torch.is_grad_enabled(), bool(x.requires_grad))

    def __enter__(self):
        self.started = self.p.begin(self.sig)

    def __exit__(self, *exc):
        if self.started:
            self.p.end()


class NLayerDiscriminator(BaseNetwork):
    """discriminator.py:61-102."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, getIntermFeat=False, stride=2):
        super().__init__()
        self.getIntermFeat, self.n_layers = getIntermFeat, n_layers
        kw = 4
        padw = int(np.ceil((kw - 1.0) / 2))
        seq = [[Conv2d(input_nc, ndf, kw, stride=stride, padding=padw), _Act()]]
        nf = ndf
        for _ in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            seq.append([norm_layer(Conv2d(nf_prev, nf, kw, stride=stride, padding=padw)), _Act()])
        nf_prev, nf = nf, min(nf * 2, 512)
        seq.append([norm_layer(Conv2d(nf_prev, nf, kw, stride=1, padding=padw)), _Act()])
        seq.append([Conv2d(nf, 1, kw, stride=1, padding=padw)])
        for n, s in enumerate(seq):
            setattr(self, 'model' + str(n), nn.Sequential(*s))

    def forward_nhwc(self, x):
        feats = []
        for n in range(self.n_layers + 2):
            first = getattr(self, 'model' + str(n))[0]
            last = n == self.n_layers + 1
            if isinstance(first, nn.Sequential):          # conv (no bias) + norm, LReLU fused into the norm kernel
                conv, norm = first
                x = norm(conv(x), act=ACT_LRELU)
            else:
                x = first(x, act=ACT_NONE if last else ACT_LRELU)
            feats.append(x)
        return feats

    def forward(self, input):
        x = ops.to_nhwc(input)
        with _planned(self, x):
            feats = [ops.nchw_view(f) for f in self.forward_nhwc(x)]
        return feats if self.getIntermFeat else feats[-1]


class MultiscaleDiscriminator(BaseNetwork):
    """discriminator.py:16-58."""

    def __init__(self, opt, input_nc, ndf=64, n_layers=3, norm_layer=None, subarch='n_layers', num_D=3,
                 getIntermFeat=False, stride=2, gpu_ids=[]):
        super().__init__()
        if subarch != 'n_layers':
            raise NotImplementedError("netD_subarch='%s' is outside the hot-path scope (SURVEY.md section 2 row 7)" % subarch)
        self.num_D, self.getIntermFeat, self.subarch = num_D, getIntermFeat, subarch
        for i in range(num_D):
            setattr(self, 'discriminator_%d' % i, NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer, getIntermFeat, stride))

    def forward(self, input, ref=None):
        return self.forward_nhwc(ops.to_nhwc(input))

    def forward_nhwc(self, x):
        """Same result for an input that is already NHWC (and possibly channel-padded, ops.pad_channels): what fsv.model packs."""
        result = []
        with _planned(self, x):
            for i in range(self.num_D):
                feats = [ops.nchw_view(f) for f in getattr(self, 'discriminator_%d' % i).forward_nhwc(x)]
                result.append(feats if self.getIntermFeat else [feats[-1]])
                if i != self.num_D - 1:
                    x = ops.avgpool3s2(x)
        return result
