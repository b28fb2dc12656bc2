"""fsv.model.Vid2VidStep (the sync-free mirror of vid2vid_model.py + loss_collector.py + face_refiner.py + input_process.py) on
CPU against the UNMODIFIED reference model: same weights, same synthetic batch, every loss of a D-step and a G-step.  The
fsv networks run on the torch emulation of fsv.ops (tests/mock_ops.py, whose pose preprocessing / face-box / crop functions are
the oracle restatements); the reference runs its own networks with `.cuda()` as a no-op.  GPU twin: tests/test_gpu_model.py."""
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
def env(monkeypatch):
    from fsv import model
    from fsv.networks import layers, generator, discriminator, vgg
    for mod in (layers, generator, discriminator, vgg, model):
        monkeypatch.setattr(mod, 'ops', mock_ops)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    return model


def _ref_model(opt, temporal):
    from models.vid2vid_model import Vid2VidModel
    m = Vid2VidModel()
    m.initialize(opt, 0)
    if opt.no_vgg_loss:
        m.lossCollector.criterionVGG = lambda a, b: 0
    if temporal:
        m.init_temporal_model()
    return m


CASES = [
    ('face', 64, 64, [], False, 1),
    ('pose', 64, 64, [], False, 1),
    ('pose', 64, 32, ['--remove_face_labels'], False, 1),
    ('street', 32, 64, [], False, 1),
    ('face', 64, 64, ['--lambda_temp', '2.0'], True, 1),
    ('pose', 64, 64, [], True, 1),
    ('face', 32, 32, ['--n_shot', '2'], False, 2),
    ('face', 64, 64, ['VGG'], False, 1),            # perceptual loss on (VGG19 with seeded random weights on both sides)
    ('pose', 128, 128, ['VGG'], False, 1),          # ... including the face-region VGG term of loss_collector.py:82
    ('face', 64, 64, ['CHUNKS'], False, 1),         # grouped spectral norm split into 3 groups (the data-parallel configuration)
]


@pytest.mark.parametrize('kind,H,W,extra,temporal,K', CASES)
def test_step_losses_match_reference_model(env, kind, H, W, extra, temporal, K):
    use_vgg = 'VGG' in extra
    if 'CHUNKS' in extra:
        from fsv.networks import layers
        env_patch = pytest.MonkeyPatch()
        env_patch.setattr(layers, 'SPECTRAL_GROUP_CHUNKS', 3)
        env_patch.setattr(layers, 'SPECTRAL_CHUNK_MIN_NUMEL', 0)
        try:
            _run_case(env, kind, H, W, [], temporal, K, False, expect_chunks=3)
        finally:
            env_patch.undo()
        return
    _run_case(env, kind, H, W, extra, temporal, K, use_vgg)


def _run_case(env, kind, H, W, extra, temporal, K, use_vgg, expect_chunks=None):
    opt = refenv.parse_opt(kind, H, W, 2, extra=TINY + [e for e in extra if e != 'VGG'], gpu=False, vgg=use_vgg)
    ref = _ref_model(opt, temporal)
    step = env.Vid2VidStep(opt)
    assert (step.vgg_loss is not None) == use_vgg
    if use_vgg:
        step.vgg_loss.vgg.load_state_dict(ref.lossCollector.criterionVGG.vgg.state_dict())
    if temporal:
        step.init_temporal_model()
    pairs = [(ref.netG, step.netG), (ref.netD, step.netD), (getattr(ref, 'netDf', None), step.netDf), (ref.netDT, step.netDT)]
    for a, b in pairs:
        assert (a is None) == (b is None)
        if a is not None:
            b.load_state_dict(a.state_dict())
    batch = synth.make(kind, 2, H, W, seed=11, K=K, temporal=temporal)
    dl = refenv.data_list(batch)
    d0 = ref(dl, mode='discriminator')
    d1 = step.discriminator_losses(batch)
    assert len(d0) == len(d1), (len(d0), list(d1))
    for (n, b), a in zip(d1.items(), d0):
        assert abs(float(a) - float(b)) < 2e-4 * max(1.0, abs(float(a))), ('D', n, float(a), float(b))
    g0, _, prev0 = ref(dl, mode='generator')
    g1, fake, prev1 = step.generator_losses(batch)
    assert list(g1) == ref.lossCollector.loss_names_G
    for (n, b), a in zip(g1.items(), g0):
        assert abs(float(a) - float(b)) < 2e-4 * max(1.0, abs(float(a))), ('G', n, [float(x) for x in g0], [float(x) for x in g1.values()])
    for a, b in zip(prev0, prev1):
        assert (a - b).abs().max() < 1e-4
    if expect_chunks is not None:
        g2, _, _ = step.generator_losses(batch)          # same call signature again: this one runs on the recorded (chunked) plan
        g3, _, _ = ref(dl, mode='generator')
        for (n, b), a in zip(g2.items(), g3):
            assert abs(float(a) - float(b)) < 2e-4 * max(1.0, abs(float(a))), ('G#2', n, float(a), float(b))
        plans = [pl for pl in step.netG.__dict__['_planner'].plans.values() if pl.get('chunks')]
        assert plans and all(len(pl['chunks']) == expect_chunks for pl in plans), [len(pl['chunks']) for pl in plans]
    sum(v.mean() for v in g1.values()).backward()
    sum(v.mean() for v in g0).backward()
    for (n, p0), (_, p1) in zip(ref.netG.named_parameters(), step.netG.named_parameters()):
        if p0.grad is not None and float(p0.grad.abs().max()) > 1e-6:
            assert p1.grad is not None, n
            e = float((p0.grad - p1.grad).norm() / p0.grad.norm())
            assert e < 5e-2, (n, e)          # relative L2 per tensor: host-wiring check (kink flips / summation order), values are checked above
