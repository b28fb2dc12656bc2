"""Two-GPU data-parallel check over NCCL (skipped on a one-GPU box): scripts/check_dp.py under torchrun -- the gradients
parallel.GradSync hands to the optimizer equal the mean over ranks of the rank-local gradients in both collect modes, and the
replayed training graph keeps the ranks' parameters identical.  (The CPU twin with gloo is tests/test_parallel_cpu.py.)"""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_gradsync_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'scripts', 'check_dp.py')]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired as e:
        pytest.fail('check_dp.py timed out; stderr tail: %s' % ((e.stderr or b'')[-3000:],))
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert p.returncode == 0 and lines, (p.stdout[-2000:], p.stderr[-3000:])
    rep = json.loads(lines[-1])
    assert rep['ok'], rep
