"""fsv: host-side mirror of the reference's network interface over the fsv_b200 CUDA kernels.

    from fsv import networks
    netG = networks.define_G(opt)           # models/networks/__init__.py:29 signature
    netD = networks.define_D(opt, ...)      # models/networks/__init__.py:41 signature

Importing this package loads few-shot-vid2vid_b200/fsv/libfsv_b200.so and fails loudly if it is
missing -- there is no PyTorch/CPU fallback on the product path.
"""
from . import _lib, ops, networks  # noqa: F401

__all__ = ['ops', 'networks']
