"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/agg_launches.py file.csv [first_id last_id]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum' or not (lo <= int(d['ID']) <= hi):
        continue
    name = re.sub(r'\(.*', '', d['Kernel Name'])
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1e6 if u in ('nsecond', 'ns') else v / 1e3 if u in ('usecond', 'us') else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005:
        continue
    print(f'{v[1]:9.3f} ms {v[0]:5d}x {100 * v[1] / tot:5.1f}%  {k[:100]}')
print(f'{tot:9.3f} ms total')
