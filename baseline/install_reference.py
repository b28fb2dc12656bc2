#!/usr/bin/env python
"""Install the UNMODIFIED reference (NVlabs/few-shot-vid2vid) under baseline/_ref so that it travels to the GPU box.

The reference is a plain Python tree without a setup.py (so `pip install /root/reference` has nothing to build); the
"install" is a verbatim copy of its importable packages (models/, util/, options/, data/) and entry scripts.  The 56 MB
`imgs/` folder (README pictures) and the FlowNet2 CUDA extensions' build products are not needed and not copied.
baseline/_ref is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored.

    python baseline/install_reference.py [--src /root/reference]

Used by: bench.py --impl reference / reference-gpu (the reference arms), tests/test_gpu_dropin.py (the reference's own
Vid2VidModel / train.py loop driving the fsv drop-in networks).  Nothing under few-shot-vid2vid_b200/ imports it.
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
KEEP = ['models', 'util', 'options', 'data', 'train.py', 'test.py', 'License.txt']


def install(src='/root/reference', quiet=False):
    if not os.path.isdir(src):
        return False
    os.makedirs(DST, exist_ok=True)
    for name in KEEP:
        s, d = os.path.join(src, name), os.path.join(DST, name)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '*.so', '*.o', 'build', '*.egg-info'))
        else:
            shutil.copy2(s, d)
    with open(os.path.join(DST, 'INSTALLED_FROM'), 'w') as f:
        f.write('%s (verbatim copy of %s)\n' % (src, ', '.join(KEEP)))
    if not quiet:
        print('reference installed under', DST)
    return True


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default=os.environ.get('FSV_REFERENCE', '/root/reference'))
    a = ap.parse_args()
    sys.exit(0 if install(a.src) else 1)
