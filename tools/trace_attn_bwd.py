"""Timeline of one CTA of the single-pass attention backward (osd_debug_attn_bwd_trace): prints, per tile, when
each softmax column group waited / worked and when the UMMA issuer served its events (SM clocks, relative)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib

B, L = 2, 8192
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 700
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
dy = torch.randn(B * L, 1024, device='cuda').to(torch.bfloat16)
y, lse = lib.attn_fwd(qkv, B, L)
lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
buf = torch.zeros(3 * 1024, dtype=torch.int64, device='cuda')
l = lib.load()
l.osd_debug_attn_bwd_trace.restype = None
l.osd_debug_attn_bwd_trace(ctypes.c_void_p(buf.data_ptr()), ctypes.c_int(cta))
lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
torch.cuda.synchronize()
l.osd_debug_attn_bwd_trace(ctypes.c_void_p(0), ctypes.c_int(0))
rec = buf.cpu().numpy().astype('uint64').reshape(3, 1024)
names = {0: 'wait_s', 1: 'got_s', 2: 'p1_done', 3: 'drained', 4: 'got_dp', 5: 'p2_done',
         10: 'PT0', 11: 'PT1', 12: 'DS0', 13: 'DS1', 14: 'S0', 15: 'S1', 16: 'DQ'}
ev = []
for slot in range(3):
    for r in rec[slot]:
        r = int(r)
        if r == 0:
            continue
        ev.append((r & 0xffffffff, slot, (r >> 48) & 0xffff, (r >> 32) & 0xffff))
ev.sort()
t0 = ev[0][0]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20, 24)
for t, slot, e, i in ev:
    if lo <= i <= hi:
        print(f'{t - t0:9d}  {"issuer" if slot == 0 else "grp" + str(slot - 1):7s} {names.get(e, e):8s} tile {i}')
# per-tile period statistics
for slot in (1, 2):
    ts = [t for t, s, e, i in ev if s == slot and e == 1]
    d = [b - a for a, b in zip(ts, ts[1:])]
    if d:
        print(f'grp{slot - 1}: tiles {len(ts)}, mean period {sum(d) / len(d):.0f} clk, min {min(d)}, max {max(d)}')
for slot, a, b, nm in ((1, 0, 1, 'wait s_full'), (1, 1, 2, 'phase 1'), (1, 2, 3, 'drain'), (1, 3, 4, 'wait dp_full'), (1, 4, 5, 'phase 2'),
                       (2, 0, 1, 'wait s_full'), (2, 1, 2, 'phase 1'), (2, 2, 3, 'drain'), (2, 3, 4, 'wait dp_full'), (2, 4, 5, 'phase 2')):
    ta = {i: t for t, s, e, i in ev if s == slot and e == a}
    tb = {i: t for t, s, e, i in ev if s == slot and e == b}
    d = [tb[i] - ta[i] for i in ta if i in tb and i >= 4]
    if d:
        print(f'grp{slot - 1} {nm:13s}: mean {sum(d) / len(d):7.0f} clk')
