"""CPU oracle for the few-shot vid2vid per-frame synthesis hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch-on-CPU tensor math,
fp32 or fp64) restatement of the reference's algorithm for the hot path
(SURVEY.md section 8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker / CPU baseline -- never as the thing measured or
shipped.  The product path (``few-shot-vid2vid_b200/fsv``) never imports it and
fails loudly when its CUDA library is missing.

Parity pin: the reference repo has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
code itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference`` with the environment shims of SURVEY.md
section 8c) and committed as fixtures under ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
from . import ops, nets  # noqa: F401
