"""Hardware checks for kernels that are NOT on the default path yet (round-2 candidates written after the round-1 GPU
budget was spent).  Skipped unless FSV_TEST_EXPERIMENTAL=1: they have not run on a B200 yet, so they must not gate the
suite; the first GPU session of round 2 runs them (under `timeout`) before anything selects those kernels by default.

    FSV_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('FSV_TEST_EXPERIMENTAL') != '1',
                                 reason='experimental kernels: set FSV_TEST_EXPERIMENTAL=1 (round-2 bring-up)')]


def test_persistent_conv_kernel_matches_default_path():
    """k_conv_tc_p (persistent CTAs, double-buffered TMEM accumulator; conv_tc.cu) against the default k_conv_tc through
    the whole tensor-core test file: the env switch is read once per process, hence the subprocess."""
    env = dict(os.environ, FSV_TC_PERSIST='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_tc.py'), '-m', 'gpu', '-q', '-x',
                        '--timeout', '120', '-p', 'no:cacheprovider'], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
