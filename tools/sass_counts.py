"""Per-kernel SASS instruction counts of libosd_b200.so (cuobjdump -sass): the mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMA* / UBLKCP = TMA, SYNCS = mbarrier) next to
the legacy ones that must be absent (HMMA = mma.sync).  usage: python tools/sass_counts.py > profiles/<tag>_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'osu-dreamer_b200', 'libosd_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
keys = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG.2D.2CTA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'UTCATOM', 'SYNCS', 'MUFU.EX2', 'FFMA2', 'HMMA', 'ATOM', 'RED']
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in txt.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0].replace('void ', '').replace('osd::', '')
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        op = m.group(1)
        counts[cur]['_all'] += 1
        for k in keys:
            if op.startswith(k) and not (k == 'UTCHMMA' and op.startswith('UTCHMMA.2CTA')):
                counts[cur][k] += 1
                total[k] += 1
print(f'{so}: {len(counts)} kernels; totals: ' + ', '.join(f'{k} {total[k]}' for k in keys))
print(f'{"kernel":58s} {"instr":>6s} ' + ' '.join(f'{k[:8]:>8s}' for k in keys))
for name, c in sorted(counts.items(), key=lambda kv: -kv[1]['UTCHMMA'] * 100000 - kv[1]['_all']):
    print(f'{name[:58]:58s} {c["_all"]:6d} ' + ' '.join(f'{c[k]:8d}' for k in keys))
