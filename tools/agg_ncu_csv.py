"""Aggregate an `ncu --csv` log (SpeedOfLight section + dram byte metrics) per kernel name: launches, total time,
DRAM bytes, achieved DRAM GB/s, and the mean of ncu's own throughput percentages.  Usage: agg_ncu_csv.py log.csv [peak_gbs]"""
import csv
import io
import json
import re
import sys
from collections import defaultdict

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
lines = open(path, errors='replace').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
per = defaultdict(lambda: defaultdict(dict))  # kernel -> id -> metric -> value
for r in rows:
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    unit = r['Metric Unit']
    m = r['Metric Name']
    if m in ('Duration', 'gpu__time_duration.sum'):
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}.get(unit, 1.0)
        m = 'ms'
    if m.startswith('dram__bytes'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    per[name][r['ID']][m] = v
out = []
for name, ids in per.items():
    n = len(ids)
    ms = sum(d.get('ms', 0.0) for d in ids.values())
    rd = sum(d.get('dram__bytes_read.sum', 0.0) for d in ids.values())
    wr = sum(d.get('dram__bytes_write.sum', 0.0) for d in ids.values())
    def mean(k):
        xs = [d[k] for d in ids.values() if k in d]
        return sum(xs) / len(xs) if xs else None
    row = {'kernel': name, 'launches': n, 'ms_total': round(ms, 3), 'ms_avg': round(ms / n, 4),
           'dram_read_GB': round(rd / 1e9, 3), 'dram_write_GB': round(wr / 1e9, 3),
           'dram_GBps': round((rd + wr) / 1e6 / ms, 1) if ms > 0 else None,
           'ncu_dram_pct': mean('DRAM Throughput'), 'ncu_compute_pct': mean('Compute (SM) Throughput'),
           'ncu_mem_pct': mean('Memory Throughput')}
    if peak:
        row['frac_of_hbm_peak'] = round(row['dram_GBps'] / peak, 3) if row['dram_GBps'] else None
    out.append(row)
out.sort(key=lambda r: -r['ms_total'])
tot = sum(r['ms_total'] for r in out)
for r in out:
    r['share'] = round(r['ms_total'] / tot, 4)
    print(f"{r['ms_total']:9.3f} ms {r['launches']:4d}x {100 * r['share']:5.1f}%  {r['dram_GBps'] or 0:7.0f} GB/s  "
          f"dram% {r['ncu_dram_pct'] or 0:5.1f}  sm% {r['ncu_compute_pct'] or 0:5.1f}  {r['kernel'][:70]}")
print(f'{tot:9.3f} ms total')
if len(sys.argv) > 3:
    json.dump({'source': path, 'hbm_peak_gbs': peak, 'kernels': out}, open(sys.argv[3], 'w'), indent=1)
