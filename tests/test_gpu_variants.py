"""Network geometries beyond the face goldens (pose-like, street-like, K = 2 attention) against the reference's own outputs, and the
non-default kernel variants (one-tile-per-CTA tcgen05 conv, per-module spectral norm, single-stream execution) through the
tensor-core / network test files in a subprocess (the switches are read once per process).  All of these ran green on the B200 in
round 2 (they were opt-in bring-up tests at the end of round 1)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('env', [{'FSV_TC_PERSIST': '0'}, {'FSV_GROUP_SPECTRAL': '0', 'FSV_WGRAD_SIDE': '0', 'FSV_BRANCH_STREAMS': '0', 'FSV_TC_BN_OCC': '0'}])
def test_non_default_kernel_variants(env):
    """FSV_TC_PERSIST=0: k_conv_tc (one tile per CTA) instead of the persistent double-buffered k_conv_tc_p; second case: per-module
    spectral norm, everything on one stream, no occupancy-aware N tile -- the round-1 execution structure."""
    files = [os.path.join(ROOT, 'tests', f) for f in ('test_gpu_tc.py', 'test_gpu_nets.py')]
    r = subprocess.run([sys.executable, '-m', 'pytest'] + files + ['-m', 'gpu', '-q', '-x', '--timeout', '300', '-p', 'no:cacheprovider'],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize('name', ['pose', 'street'])
def test_generator_other_geometries_vs_reference_golden(name):
    """The drop-in generator in the pose-like (6-channel, portrait H = 2W, warp + spade_combine) and street-like (wide
    W = 2H, no flow branch) configurations against the reference's own outputs (tests/golden/g_variants_tiny.npz, pinned
    for the oracle on CPU in test_oracle_golden.py)."""
    import json
    from argparse import Namespace
    import torch
    from fsv import networks, ops
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fsvtest import load_npz, state_from, T, rel_err, l2_err
    z = load_npz('g_variants_tiny.npz')
    pre = name + '.'
    opt = Namespace(**json.loads(str(z[pre + 'opt'])))
    opt.gpu_ids = [0]
    old = ops.CONV_USE_TC
    ops.CONV_USE_TC = 0
    try:
        G = networks.define_G(opt)
        G.load_state_dict(state_from(z, pre + 'sd.'))
        G.train()
        label = T(z[pre + 'label']).cuda().requires_grad_(True)
        out = G(label, T(z[pre + 'lref']).cuda(), T(z[pre + 'iref']).cuda())
        assert rel_err(out[0], T(z[pre + 'out_img'])) < 1e-3
        loss = (out[0] * T(z[pre + 'r1']).cuda()).sum()
        if int(z[pre + 'has_flow']):
            assert rel_err(out[1][0], T(z[pre + 'out_flow'])) < 1e-3
            assert rel_err(out[2][0], T(z[pre + 'out_mask'])) < 1e-3
            loss = loss + 0.05 * out[1][0].sum() + out[2][0].sum()
        else:
            assert out[1][0] is None and out[2][0] is None
        loss.backward()
        assert l2_err(label.grad, T(z[pre + 'grad_label'])) < 1e-2
        params = dict(G.named_parameters())
        for k in z.files:
            if k.startswith(pre + 'grad.') and k != pre + 'grad_label':
                assert l2_err(params[k[len(pre) + 5:]].grad, T(z[k])) < 1e-2, k
    finally:
        ops.CONV_USE_TC = old


def test_generator_two_reference_images_vs_reference_golden():
    """K = 2 attention path of the drop-in generator (attention GEMMs as per-sample 1x1 convs + channel softmax of the
    C ABI) against the reference (tests/golden/g_kshot_tiny.npz; the oracle is pinned to the same file on CPU).
    """
    import torch
    from fsv import networks, ops
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fsvtest import load_npz, state_from, opt_from, T, rel_err, l2_err
    z = load_npz('g_kshot_tiny.npz')
    opt = opt_from(z)
    opt.gpu_ids = [0]
    old = ops.CONV_USE_TC
    ops.CONV_USE_TC = 0
    try:
        G = networks.define_G(opt)
        G.load_state_dict(state_from(z, 'sd.'))
        G.train()
        label = T(z['label']).cuda().requires_grad_(True)
        out = G(label, T(z['lref']).cuda(), T(z['iref']).cuda())
        assert rel_err(out[0], T(z['out_img'])) < 1e-3
        assert rel_err(out[1][0], T(z['out_flow'])) < 1e-3
        assert rel_err(out[2][0], T(z['out_mask'])) < 1e-3
        assert rel_err(out[4][0], T(z['out_warp'])) < 1e-3
        assert rel_err(out[7], T(z['atn_vis'])) < 1e-3
        assert torch.equal(out[8].cpu(), torch.from_numpy(z['ref_idx']))
        loss = (out[0] * T(z['r1']).cuda()).sum() + 0.05 * out[1][0].sum() + out[2][0].sum()
        loss.backward()
        assert l2_err(label.grad, T(z['grad_label'])) < 1e-2
        params = dict(G.named_parameters())
        for k in z.files:
            if k.startswith('grad.'):
                assert l2_err(params[k[5:]].grad, T(z[k])) < 1e-2, k
    finally:
        ops.CONV_USE_TC = old


def test_generator_two_reference_images_chunked_attention_on_gpu():
    """No-grad K = 2 forward with the attention matrix formed a few query rows at a time (memory-bounded form of the inference sweep)
    against the one-piece form on the same kernels: frame, flow, warp, attention visualisation, picked reference."""
    import torch
    from fsv import networks, ops
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fsvtest import load_npz, state_from, opt_from, T, rel_err
    z = load_npz('g_kshot_tiny.npz')
    opt = opt_from(z)
    opt.gpu_ids = [0]
    old = ops.CONV_USE_TC
    ops.CONV_USE_TC = 0
    try:
        G = networks.define_G(opt)
        sd = state_from(z, 'sd.')
        G.train()
        label, lref, iref = T(z['label']).cuda(), T(z['lref']).cuda(), T(z['iref']).cuda()
        outs = []
        with torch.no_grad():
            for budget in (1 << 40, 1, 3 * 4 * label.shape[0] * (label.shape[3] >> G.n_downsample_A) * G.n_shot *
                           (label.shape[2] >> G.n_downsample_A) * (label.shape[3] >> G.n_downsample_A)):     # one piece; 1 row; 3 rows per chunk
                G.load_state_dict(sd)
                G.attention_chunk_bytes = budget
                outs.append(G(label, lref, iref))
        a = outs[0]
        for b in outs[1:]:
            assert rel_err(b[0], a[0]) < 1e-5 and rel_err(b[1][0], a[1][0]) < 1e-5 and rel_err(b[4][0], a[4][0]) < 1e-5
            assert rel_err(b[7], a[7]) < 1e-5 and torch.equal(a[8], b[8])
    finally:
        ops.CONV_USE_TC = old
