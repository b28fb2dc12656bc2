#!/usr/bin/env python
"""Reads the .jsonl of scripts/trace_step.py: concurrency histogram, time each kernel runs ALONE (the serial part of the step), a per-millisecond
view of what is running.  usage: scripts/analyze_trace.py gpurun_out/trace.jsonl [--ms]"""
import collections
import json
import sys


def main():
    ev = [json.loads(l) for l in open(sys.argv[1])]
    pts = []
    for i, e in enumerate(ev):
        pts.append((e['ts_us'], 1, i))
        pts.append((e['ts_us'] + e['dur_us'], -1, i))
    pts.sort()
    active, last = set(), 0
    alone, hist, total = collections.defaultdict(float), collections.Counter(), collections.defaultdict(lambda: [0.0, 0])
    for t, d, i in pts:
        hist[len(active)] += t - last
        if len(active) == 1:
            alone[ev[next(iter(active))]['name'].split('(')[0]] += t - last
        last = t
        if d == 1:
            active.add(i)
        else:
            active.discard(i)
    for e in ev:
        k = total[e['name'].split('(')[0]]
        k[0] += e['dur_us']
        k[1] += 1
    print('wall %.1f us; kernels running -> us: %s' % (last, {k: round(v) for k, v in sorted(hist.items())}))
    print('--- time alone / total / launches, by kernel')
    for n, d in sorted(alone.items(), key=lambda kv: -kv[1])[:28]:
        print('%9.1f %9.1f x%-4d %s' % (d, total[n][0], total[n][1], n[:100]))
    if '--ms' in sys.argv:
        B = 1000.0
        nb = int(last // B) + 1
        buck = [collections.defaultdict(float) for _ in range(nb)]
        for e in ev:
            s, t = e['ts_us'], e['ts_us'] + e['dur_us']
            n = e['name'].split('(')[0].replace('void ', '')[:24] + '@%s' % e['stream']
            b = int(s // B)
            while s < t:
                nxt = min(t, (b + 1) * B)
                buck[b][n] += nxt - s
                s = nxt
                b += 1
        for i, b in enumerate(buck):
            top = sorted(b.items(), key=lambda kv: -kv[1])[:4]
            print('%2d ms busy %5.0f | ' % (i, sum(b.values())) + '  '.join('%s %.0f' % (k, v) for k, v in top))


if __name__ == '__main__':
    main()
