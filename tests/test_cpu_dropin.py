"""INTEGRATION.md option B on CPU: the reference's own Vid2VidModel / LossCollector / FaceRefineModel (UNMODIFIED, imported
from baseline/_ref or /root/reference, `.cuda()` turned into a no-op) driving `fsv.networks.define_G/define_D`, whose
`fsv.ops` is replaced by the torch-CPU emulation of tests/mock_ops.py (test infrastructure; the product has no CPU path).
Checks the host-side contract of the drop-in for the three dataset geometries of BASELINE.json -- constructor arguments the
reference passes (D input channels 8 / 20 / 46, the face discriminator), state_dict interchange, the 9-tuple consumed by
generate_images, 5-D reshapes, D feature lists consumed by GANLoss / feature matching -- by comparing every loss of a
D-step and a G-step with the reference's own networks on the same weights and inputs.  The GPU twin is tests/test_gpu_dropin.py."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import refenv   # noqa: E402
import synth    # noqa: E402
import mock_ops  # noqa: E402

pytestmark = pytest.mark.skipif(not refenv.available(), reason='reference not installed (baseline/_ref) and /root/reference absent')

TINY = ['--ngf', '4', '--nff', '4', '--ndf', '4', '--n_downsample_G', '3', '--n_adaptive_layers', '2', '--n_downsample_F', '2', '--n_blocks_F', '1']


@pytest.fixture()
def cpu_reference(monkeypatch):
    from fsv import networks
    from fsv.networks import layers, generator, discriminator
    for mod in (layers, generator, discriminator):
        monkeypatch.setattr(mod, 'ops', mock_ops)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    yield networks
    refenv.unpatch_networks()


def _model(opt, networks=None):
    from models.vid2vid_model import Vid2VidModel
    if networks is not None:
        refenv.patch_networks(networks.define_G, networks.define_D)
    try:
        m = Vid2VidModel()
        m.initialize(opt, 0)
    finally:
        refenv.unpatch_networks()
    m.lossCollector.criterionVGG = lambda a, b: 0
    return m


@pytest.mark.parametrize('kind,H,W,extra', [('face', 64, 64, []), ('pose', 64, 64, []), ('pose', 64, 32, ['--remove_face_labels']),
                                            ('street', 32, 64, [])])
def test_reference_model_on_dropin_networks(cpu_reference, kind, H, W, extra):
    opt = refenv.parse_opt(kind, H, W, 2, extra=TINY + extra, gpu=False)
    ref = _model(opt)
    mine = _model(opt, cpu_reference)
    for n in ('netG', 'netD', 'netDf'):
        a, b = getattr(ref, n, None), getattr(mine, n, None)
        assert (a is None) == (b is None)
        if a is not None:
            assert type(b).__module__.startswith('fsv.')
            b.load_state_dict(a.state_dict())
    dl = refenv.data_list(synth.make(kind, 2, H, W, seed=3))
    for mode in ('discriminator', 'generator'):
        l0 = ref(dl, mode=mode)
        l1 = mine(dl, mode=mode)
        if mode == 'generator':
            l0, l1 = l0[0], l1[0]
        assert len(l0) == len(l1)
        for a, b in zip(l0, l1):
            assert abs(float(a) - float(b)) < 2e-4 * max(1.0, abs(float(a))), (kind, mode, [float(x) for x in l0], [float(x) for x in l1])
        sum(x.mean() for x in l1).backward()
    assert any(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in mine.netG.parameters())
