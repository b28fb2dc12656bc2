"""CPU checks of the measurement tooling: the timeline analysis (scripts/analyze_trace.py) on a synthetic kernel record, and the join of the committed
ncu capture with the bench line (bench.ncu_traffic -> roofline.traffic)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_analyze_trace_on_a_synthetic_timeline(tmp_path):
    """two streams: kernel a runs alone for 10 us, then overlaps b for 5 us, b alone for 15 us, a gap of 2 us, c alone for 8 us"""
    ev = [dict(name='k_a(int)', ts_us=0.0, dur_us=15.0, stream=7, cat='kernel', grid=[1, 1, 1]),
          dict(name='k_b(float*)', ts_us=10.0, dur_us=20.0, stream=9, cat='kernel', grid=[2, 1, 1]),
          dict(name='k_c()', ts_us=32.0, dur_us=8.0, stream=7, cat='kernel', grid=[3, 1, 1])]
    p = tmp_path / 't.jsonl'
    p.write_text('\n'.join(json.dumps(e) for e in ev) + '\n')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'analyze_trace.py'), str(p), '--ms'], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0].startswith('wall 40.0 us')
    assert '{0: 2, 1: 33, 2: 5}' in lines[0]            # 2 us idle, 33 us with one kernel, 5 us with two
    alone = {l.split()[-1]: float(l.split()[0]) for l in lines[2:5]}
    assert alone == {'k_a': 10.0, 'k_b': 15.0, 'k_c': 8.0}


def test_bench_reads_the_committed_ncu_traffic():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_traffic()
    assert set(t) == {'conv', 'spade'}
    for k in t.values():
        # DRAM traffic of the captured launch does not exceed its algorithmic bytes by more than 20 % (no wasted re-reads)
        assert 0 < k['dram_bytes_per_launch'] <= 1.2 * k['algorithmic_bytes_per_launch']
        assert k['launch_us_under_ncu'] > 0 and 'k_' in k['layer']
