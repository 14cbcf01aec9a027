"""Condensed view of an `ncu --page source --csv --print-source sass` export: key instructions + sample shares.
usage: python tools/sass_profile.py export.csv [min_pct]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[idx['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
key = re.compile(r'SYNCS|UTCHMMA|LDTM|STTM|UTMA|BAR|UTCBAR|EXIT|FENCE|LDG|STG|UTCATOM|ARRIVE|MEMBAR|UBLKCP')
acc, accst, first, mufu = 0, {}, None, 0


def flush():
    global acc, accst, first, mufu
    if acc > tot * minp / 100:
        top = sorted(((v, k[6:]) for k, v in accst.items()), reverse=True)[:3]
        print(f'{first:5d}+ {acc:6d} {100 * acc / tot:4.1f}%  ({mufu} MUFU) {top}')
    acc, accst, first, mufu = 0, {}, None, 0


print('total samples', tot)
for n, r in enumerate(data):
    s = int(r[idx['# Samples']])
    src = r[idx['Source']].strip()
    if key.search(src):
        flush()
        top = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        if s > tot * minp / 100 or 'SYNCS.PHASECHK.TRANS64.TRYWAIT' in src or 'UTCHMMA' in src or 'LDTM' in src:
            print(f'{n:5d}  {s:6d} {100 * s / tot:4.1f}% x{r[idx["Instructions Executed"]]:>9} {src[:60]:60s} {top if s > tot * 0.002 else ""}')
    else:
        if first is None:
            first = n
        acc += s
        mufu += 'MUFU' in src
        for h in stalls:
            accst[h] = accst.get(h, 0) + int(r[idx[h]])
flush()
