#!/usr/bin/env python
"""gpurun_out/prof_layers_r2.ncu-rep (ncu --set full of scripts/ncu_layers.py) + gpurun_out/ncu_layers.json (algorithmic bytes of the
layers it runs) -> profiles/ncu_layers_r2_summary.txt and profiles/ncu_traffic_r2.json (read by bench.py for `roofline.traffic`).
Run in the build container (ncu reads the report offline)."""
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic']


def to_bytes(v, unit):
    return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_us(v, unit):
    return float(v.replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit, 1)


def main():
    rep = 'gpurun_out/prof_layers_r2.ncu-rep'
    info = json.load(open('gpurun_out/ncu_layers.json'))
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units = r[0], r[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = ['ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_p|k_spade_tc|k_wgrad_tc_mn python scripts/ncu_layers.py',
           '(B200, final round-2 code; layers of known shape: %s)' % '; '.join('%s = %s' % (k, v['shape']) for k, v in info.items()),
           'per launch: duration, DRAM bytes read + written (cold-cache replay under ncu) against the algorithmic bytes of the layer']
    rows = []
    for row in r[2:]:
        if len(row) < len(hdr):
            continue
        name = row[col['Kernel Name']].split('(')[0]
        rec = dict(kernel=name, grid=row[col['Grid Size']], us=to_us(row[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']]),
                   dram=to_bytes(row[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']]) +
                   to_bytes(row[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']]))
        for k in WANT[6:]:
            if k in col:
                rec[k] = row[col[k]]
        rows.append(rec)
        out.append('  %-40s grid %-16s %8.1f us  dram %8.2f MB  tensor %s%%  lts %s%%  dram %s%%  regs %s' % (
            name[:40], rec['grid'], rec['us'], rec['dram'] / 1e6, rec.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', '?')[:5],
            rec.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', '?')[:5], rec.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '?')[:5],
            rec.get('launch__registers_per_thread', '?')))
    traffic = {'layers': info, 'launches': rows}
    open('profiles/ncu_layers_r2_summary.txt', 'w').write('\n'.join(out) + '\n')
    json.dump(traffic, open('profiles/ncu_traffic_r2.json', 'w'), indent=1)
    print('\n'.join(out))


if __name__ == '__main__':
    main()
