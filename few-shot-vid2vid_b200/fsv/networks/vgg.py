"""VGG19 feature stack of the perceptual loss on the fsv kernels.

Drop-in for ``VGG_Activations`` of the reference (models/networks/vgg.py:45-59: ``torchvision.models.vgg19().features`` as a
ModuleList, the activations after the layers ``feature_idx`` = [1, 6, 11, 20, 29] = relu1_1 ... relu5_1) and for ``VGGLoss``
(models/networks/loss.py:107-128: sum_i w_i L1(vgg(x)_i, vgg(y)_i.detach()), w = 1/32, 1/16, 1/8, 1/4, 1).  Same ``state_dict`` keys
(``features.{0,2,5,...}.weight / bias``), so the torchvision ImageNet checkpoint loads unchanged where it is available; offline it
is not, so tests and benchmarks run the same architecture with seeded random weights on both sides.  Conv + ReLU pairs are one
fused kernel call (3x3 tcgen05 / thin-input conv with a ReLU epilogue); layers beyond the last requested index do not influence
the result and are not evaluated (the reference runs all 37).
"""
import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_RELU
from .layers import Conv2d

CFG_E = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M']


class _Slot(nn.Module):
    """parameter-free position of the torchvision Sequential (ReLU / MaxPool2d): keeps the layer indices"""

    def __init__(self, kind):
        super().__init__()
        self.kind = kind


class VGGActivations(nn.Module):
    def __init__(self, feature_idx=(1, 6, 11, 20, 29)):
        super().__init__()
        layers, cin = [], 3
        for v in CFG_E:
            if v == 'M':
                layers.append(_Slot('pool'))
            else:
                layers += [Conv2d(cin, v, 3, padding=1), _Slot('relu')]
                cin = v
        self.features = nn.ModuleList(layers)
        self.idx_list = list(feature_idx)
        for p in self.parameters():
            p.requires_grad = False            # loss.py:110 uses the network as a fixed feature extractor

    def forward(self, x):
        """x: NCHW (B, 3, H, W) -> list of NCHW-shaped activations at ``idx_list``."""
        h = ops.to_nhwc(x)
        out, last = [], max(self.idx_list)
        i = 0
        while i <= last:
            m = self.features[i]
            if isinstance(m, Conv2d):
                h = m(h, act=ACT_RELU)         # conv + the in-place ReLU that follows it (index i + 1)
                if i in self.idx_list:
                    raise NotImplementedError('pre-ReLU activations are not exposed (the reference asks for post-ReLU indices)')
                i += 1
            elif m.kind == 'pool':
                h = ops.maxpool2(h)
            if i in self.idx_list:
                out.append(ops.nchw_view(h))
            i += 1
        return out


class VGGLoss(nn.Module):
    """loss.py:107-128."""

    def __init__(self):
        super().__init__()
        self.vgg = VGGActivations()
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]

    def forward(self, x, y):
        if x.dim() == 5:
            x, y = x.reshape(-1, *x.shape[2:]), y.reshape(-1, *y.shape[2:])
        with torch.no_grad():
            y_vgg = self.vgg(y)
        x_vgg = self.vgg(x)
        loss = 0
        for w, a, b in zip(self.weights, x_vgg, y_vgg):
            loss = loss + w * torch.nn.functional.l1_loss(a, b.detach())
        return loss
