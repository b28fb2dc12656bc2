"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/fsv_b200.h declares; the drop-in modules expose exactly the reference's state_dict; the
product refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from fsvtest import load_npz, state_from, opt_from

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'fsv_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(fsv_[a-z0-9_]+)\s*\(', txt)))


def test_library_loads_and_exports_every_declared_symbol():
    from fsv import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), 'libfsv_b200.so does not export %s' % n
    # and the ctypes table binds each of them (except the error string getter bound separately)
    for n in names:
        assert n in _lib.SIGNATURES or n == 'fsv_last_error', n
    assert lib.fsv_version() >= 100


def test_state_dict_names_match_reference():
    from fsv import networks
    z = load_npz('g_face_tiny.npz')
    opt = opt_from(z)
    G = networks.define_G(opt)
    ref = state_from(z, 'sd.')
    mine = G.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    assert all(tuple(mine[k].shape) == tuple(ref[k].shape) for k in ref)
    G.load_state_dict(ref)
    G.init_temporal_network()
    zt = load_npz('g_face_tiny_temporal.npz')
    assert set(G.state_dict().keys()) == set(state_from(zt, 'sd.').keys())
    zd = load_npz('d_tiny.npz')
    D = networks.define_D(opt, 8, opt.ndf, opt.n_layers_D, opt.norm_D, opt.netD_subarch, 2, True, gpu_ids=[])
    assert list(D.state_dict().keys()) == list(state_from(zd, 'sd.').keys())
    # finetune's name filter (vid2vid_model.py:208-209) still finds its parameter groups
    keys = list(mine.keys())
    assert any('fc' in k for k in keys) and any('conv_img' in k for k in keys) and any(k.startswith('up_') for k in keys)


@pytest.mark.parametrize('name', ['pose', 'street'])
def test_state_dict_names_match_reference_other_geometries(name):
    """Same check on the pose-like (6-channel, H = 2W, warp + spade_combine) and street-like (W = 2H, no flow branch)
    configurations: the drop-in modules build the reference's parameter / buffer set key for key."""
    import json
    from argparse import Namespace
    from fsv import networks
    z = load_npz('g_variants_tiny.npz')
    opt = Namespace(**json.loads(str(z[name + '.opt'])))
    G = networks.define_G(opt)
    ref = state_from(z, name + '.sd.')
    mine = G.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    assert all(tuple(mine[k].shape) == tuple(ref[k].shape) for k in ref)
    G.load_state_dict(ref)


def test_state_dict_names_match_reference_kshot():
    """n_shot = 2: the attention key / query encoders are built with the reference's names, in its registration order."""
    from fsv import networks
    z = load_npz('g_kshot_tiny.npz')
    G = networks.define_G(opt_from(z))
    ref = state_from(z, 'sd.')
    mine = G.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    assert all(tuple(mine[k].shape) == tuple(ref[k].shape) for k in ref)
    assert any(k.startswith('atn_key_first.') for k in mine) and any(k.startswith('atn_query_1.') for k in mine)
    G.load_state_dict(ref)


def test_init_matches_reference_statistics():
    from fsv import networks
    z = load_npz('g_face_tiny.npz')
    opt = opt_from(z)
    torch.manual_seed(0)
    G = networks.define_G(opt)
    ref = state_from(z, 'sd.')
    mine = G.state_dict()
    for k in ('up_2.conv_0.weight_orig', 'fc_spade_0_1.0.weight_orig', 'label_embedding.down_1.0.weight'):
        assert abs(mine[k].std().item() / ref[k].std().item() - 1) < 0.25, k
    for k in mine:
        if k.endswith('.bias'):
            assert float(mine[k].abs().max()) == 0.0, k
        if k.endswith('bn.weight'):
            assert float((mine[k] - 1).abs().max()) == 0.0, k


def test_product_has_no_cpu_fallback():
    from fsv import networks
    z = load_npz('g_face_tiny.npz')
    opt = opt_from(z)
    G = networks.define_G(opt)
    with pytest.raises(Exception) as e:
        G(torch.zeros(1, 1, 64, 64), torch.zeros(1, 1, 1, 64, 64), torch.zeros(1, 1, 3, 64, 64))
    assert 'CUDA' in str(e.value) or 'cuda' in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'few-shot-vid2vid_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(d, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, os.path.join(d, f)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm: the unmodified reference from baseline/_ref on host threads) keeps the driver's
    output contract: exactly one JSON line on stdout with the metric keys, and it must not load this repo's product;
    the fsv arm refuses to run without a CUDA device (no CPU fallback)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--workload', 'tiny'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl',) if 'unavailable' in d else ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'dtype', 'data', 'config', 'e2e', 'cpu_baseline'):
        assert k in d, k
    if 'unavailable' in d:
        assert not os.path.isdir(os.path.join(root, 'baseline', '_ref'))
        return
    assert d['impl'] == 'reference' and d['value'] > 0 and d['cpu_baseline']['kind'] == 'reference'
    src = open(os.path.join(root, 'baseline', 'ref_arm.py')).read() + open(os.path.join(root, 'baseline', 'refenv.py')).read()
    assert 'import fsv' not in src and 'from fsv' not in src and 'import oracle' not in src
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    if not torch.cuda.is_available():
        r2 = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '1', '--warmup', '0'],
                            capture_output=True, text=True, timeout=600)
        assert r2.returncode != 0 and 'no CPU fallback' in (r2.stderr + r2.stdout)


def test_bench_options_equal_the_reference_parser():
    """bench.make_opt (the fsv arm builds its options without the reference installed) must agree with what the reference's own
    option parser produces for the same flags (the reference arms use that parser) on every field both define."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'baseline'))
    import refenv
    if not refenv.available():
        import pytest
        pytest.skip('reference not installed')
    import bench
    for name, wl in bench.WORKLOADS.items():
        mine = vars(bench.make_opt(name))
        ref = vars(refenv.parse_opt(wl['kind'], wl['H'], wl['W'], wl['batch'], extra=wl.get('ref_extra', []), gpu=False, vgg=bool(wl.get('vgg'))))
        for k, v in mine.items():
            if k in ref and k not in ('gpu_ids', 'for_face'):
                assert ref[k] == v, (name, k, ref[k], v)
