"""Factory functions with the reference's signatures (models/networks/__init__.py:29-55)."""
import torch

from .generator import FewShotGenerator, FlowGenerator, LabelEmbedder, BaseNetwork  # noqa: F401
from .discriminator import MultiscaleDiscriminator, NLayerDiscriminator, get_nonspade_norm_layer  # noqa: F401


def define_G(opt):
    if 'fewshot' in opt.netG:
        netG = FewShotGenerator(opt)
    else:
        raise ValueError('generator not implemented!')
    if opt.isTrain and getattr(opt, 'print_G', False):
        netG.print_network()
    moved = len(opt.gpu_ids) > 0
    if moved:
        assert torch.cuda.is_available()
        netG.cuda()
    # after the move the reference's init no longer reaches weight_orig of spectral layers (see layers.init_weights)
    netG.init_weights(opt.init_type, opt.init_variance, reach_spectral=not moved)
    return netG


def define_D(opt, input_nc, ndf, n_layers_D, norm='spectralinstance', subarch='n_layers', num_D=1, getIntermFeat=False,
             stride=2, gpu_ids=[]):
    norm_layer = get_nonspade_norm_layer(opt, norm_type=norm)
    if opt.which_model_netD == 'multiscale':
        netD = MultiscaleDiscriminator(opt, input_nc, ndf, n_layers_D, norm_layer, subarch, num_D, getIntermFeat, stride, gpu_ids)
    elif opt.which_model_netD == 'n_layers':
        netD = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer, getIntermFeat)
    else:
        raise ValueError('unknown type discriminator %s!' % opt.which_model_netD)
    if opt.isTrain and getattr(opt, 'print_D', False):
        netD.print_network()
    moved = len(gpu_ids) > 0
    if moved:
        assert torch.cuda.is_available()
        netD.cuda()
    netD.init_weights(opt.init_type, opt.init_variance, reach_spectral=not moved)
    return netD
