"""Run the UNMODIFIED reference (baseline/_ref, see install_reference.py) on this software stack.

Environment shims of SURVEY.md section 8c -- none of them edits a reference file:
  * `apex.parallel.SyncBatchNorm` -> torch.nn.SyncBatchNorm (normalization.py:15,33,80; apex is not installable here),
  * a stub `dominate` module (util/html.py:8), `fractions.gcd = math.gcd` (models/trainer.py:14,32),
  * `torch.optim.Adam` accepting the reference's betas=(0, 0.999) int/float mix (base_model.py:45-48),
  * with `--no_vgg_loss` the reference's `LossCollector.discriminate_face` still calls `self.criterionVGG`
    (loss_collector.py:82, an AttributeError in the reference): `criterionVGG` is bound to a zero function, which is what
    `--no_vgg_loss` means everywhere else in that file (loss_collector.py:122-129),
  * `torchvision.models.vgg19(pretrained=True)` (models/networks/vgg.py:16,48; a download) -> the same network with seeded random
    weights, for runs that keep the VGG loss on,
  * CPU only (tests/golden generation in the build container): device-agnostic `resample` and ByteTensor aliases.

This module imports neither fsv nor oracle.  `patch_networks(define_G, define_D)` is INTEGRATION.md option B: the
reference's own Vid2VidModel / LossCollector / train.py loop then drive whatever factories are passed in.
"""
import fractions
import math
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_root():
    for cand in (os.environ.get('FSV_REFERENCE'), os.path.join(HERE, '_ref'), '/root/reference'):
        if cand and os.path.isdir(os.path.join(cand, 'models')):
            return cand
    return None


def available():
    return ref_root() is not None


def install_shims(cpu=False):
    root = ref_root()
    if root is None:
        raise RuntimeError('reference not installed: run `python baseline/install_reference.py` where /root/reference exists')
    if 'apex' not in sys.modules:
        apex = types.ModuleType('apex')
        par = types.ModuleType('apex.parallel')
        par.SyncBatchNorm = torch.nn.SyncBatchNorm
        apex.parallel = par
        sys.modules['apex'], sys.modules['apex.parallel'] = apex, par
    if 'dominate' not in sys.modules:
        dom = types.ModuleType('dominate')
        dom.tags = types.ModuleType('dominate.tags')
        sys.modules['dominate'], sys.modules['dominate.tags'] = dom, dom.tags
    fractions.gcd = math.gcd
    if not getattr(torch.optim.Adam, '_fsv_float_betas', False):
        class Adam(torch.optim.Adam):
            # base_model.py:45 passes betas=(0, 0.999): torch >= 2.x rejects the int/float mix the reference was written against
            _fsv_float_betas = True

            def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), **kw):
                super().__init__(params, lr=lr, betas=(float(betas[0]), float(betas[1])), **kw)
        torch.optim.Adam = Adam
    if root not in sys.path:
        sys.path.insert(0, root)
    _shim_vgg()
    if cpu:
        import torch.nn.functional as F

        def resample_any_device(image, flow):
            # base_network.py:28-37 without the hard .cuda() calls (same arithmetic)
            b, c, h, w = image.shape
            hor = torch.linspace(-1.0, 1.0, w).view(1, 1, 1, w).expand(b, 1, h, w)
            ver = torch.linspace(-1.0, 1.0, h).view(1, 1, h, 1).expand(b, 1, h, w)
            grid = torch.cat([hor, ver], 1).to(flow.device)
            flow = torch.cat([flow[:, 0:1] / ((w - 1.0) / 2.0), flow[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
            return F.grid_sample(image, (grid + flow).permute(0, 2, 3, 1), mode='bilinear', padding_mode='border', align_corners=True)
        import models.networks.generator as refgen
        import models.loss_collector as reflc
        import models.input_process as refip
        refgen.resample = resample_any_device
        reflc.resample = resample_any_device
        refip.torch = _TorchCudaProxy(torch)
    return root


def _shim_vgg():
    """models/networks/vgg.py:16,48 ask torchvision for ImageNet weights (a download; there is no network here): hand back the same
    VGG19 with seeded random weights -- same architecture and cost, used identically by every arm / test that enables the VGG loss."""
    try:
        import torchvision
    except Exception:
        return
    if getattr(torchvision.models.vgg19, '_fsv_offline', False):
        return
    real = torchvision.models.vgg19

    def vgg19(pretrained=False, **kw):
        kw.pop('weights', None)
        state = torch.get_rng_state()
        torch.manual_seed(1919)
        try:
            return real(weights=None, **kw)
        finally:
            torch.set_rng_state(state)
    vgg19._fsv_offline = True
    torchvision.models.vgg19 = vgg19


class _TorchCudaProxy:
    """`torch` as seen by models/input_process.py on a CPU-only box: torch.cuda.ByteTensor / FloatTensor construct CPU tensors."""

    def __init__(self, real):
        self._real = real

        def dims(s):
            return [int(x) for x in (s[0] if len(s) == 1 and isinstance(s[0], (tuple, list, real.Size)) else s)]
        self.cuda = types.SimpleNamespace(ByteTensor=lambda *s: real.zeros(*dims(s), dtype=real.uint8),
                                          FloatTensor=lambda *s: real.zeros(*dims(s)))

    def __getattr__(self, k):
        return getattr(self._real, k)


DATASET_FLAGS = {
    'face': ['--dataset_mode', 'fewshot_face', '--adaptive_spade', '--warp_ref', '--spade_combine'],
    'pose': ['--dataset_mode', 'fewshot_pose', '--adaptive_spade', '--warp_ref', '--spade_combine', '--add_face_D'],
    'street': ['--dataset_mode', 'fewshot_street', '--adaptive_spade'],
}


def parse_opt(kind, H, W, batch, extra=(), gpu=True, train=True, ckpt='/tmp/fsv_ref_ckpt', vgg=False):
    """The reference's own option parser (options/train_options.py) on the BASELINE.json flags of a dataset kind.
    fineSize is the frame WIDTH, aspect_ratio = W / H (fewshot_pose_dataset.py:100, generator.py:83-84)."""
    install_shims(cpu=not gpu)
    argv = list(DATASET_FLAGS[kind]) + ['--no_flow_gt'] + ([] if vgg else ['--no_vgg_loss']) + ['--loadSize', str(W), '--fineSize', str(W),
                                        '--aspect_ratio', repr(W / H), '--batchSize', str(batch), '--checkpoints_dir', ckpt,
                                        '--gpu_ids', '0' if gpu else '-1', '--name', 'fsv_%s_%dx%d' % (kind, H, W)] + list(extra)
    old = sys.argv
    sys.argv = ['train.py'] + argv
    try:
        if train:
            from options.train_options import TrainOptions
            opt = TrainOptions().parse()
        else:
            from options.test_options import TestOptions
            opt = TestOptions().parse()
    finally:
        sys.argv = old
    return opt


_ORIG = {}


def patch_networks(define_G, define_D):
    """INTEGRATION.md option B: rebind the two factory functions of models/networks/__init__.py:29-55."""
    import models.networks as refnets
    _ORIG.setdefault('G', refnets.define_G)
    _ORIG.setdefault('D', refnets.define_D)
    refnets.define_G, refnets.define_D = define_G, define_D


def unpatch_networks():
    import models.networks as refnets
    if _ORIG:
        refnets.define_G, refnets.define_D = _ORIG['G'], _ORIG['D']


def create_model(opt):
    """models/models.py:16-38 create_model + the no-VGG shim for --add_face_D.  -> (model, optimizer_G, optimizer_D)"""
    from models.models import create_model as ref_create
    model, flow_net, (opt_g, opt_d) = ref_create(opt, 0)
    lc = model.module.lossCollector
    if opt.no_vgg_loss and not hasattr(lc, 'criterionVGG'):
        lc.criterionVGG = lambda a, b: 0
    return model, opt_g, opt_d


def data_list(batch, device=None):
    """train.py:44-52: [tgt_label, tgt_image, flow_gt, conf_gt] + [ref_label, ref_image] + data_prev for one frame (t = 0)."""
    mv = (lambda t: t) if device is None else (lambda t: t.to(device, non_blocking=True))
    prev = [None, None, None]
    if 'prev_label' in batch:
        prev = [mv(batch['prev_label']), mv(batch['prev_real']), mv(batch['prev_fake'])]
    return [mv(batch['tgt_label']), mv(batch['tgt_image']), [None, None], [None, None], mv(batch['ref_label']), mv(batch['ref_image'])] + prev


def train_iteration(opt, model, opt_g, opt_d, dl):
    """train.py:55-62 inner loop body for one frame.  Returns (d_losses, g_losses) as the reference's lists."""
    from models.loss_collector import loss_backward
    d_losses = model(dl, mode='discriminator')
    d_losses = loss_backward(opt, d_losses, opt_d, 1)
    g_losses, generated, prev = model(dl, save_images=False, mode='generator')
    g_losses = loss_backward(opt, g_losses, opt_g, 0)
    return d_losses, g_losses
