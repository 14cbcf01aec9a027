"""Timeline of one CTA of the forward attention kernel variant 5 (osd_debug_attn_fwd_trace)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib

B, L = 2, 8192
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 700
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=5)
buf = torch.zeros(2 * 1024, dtype=torch.int64, device='cuda')
l = lib.load()
l.osd_debug_attn_fwd_trace.restype = None
l.osd_debug_attn_fwd_trace(ctypes.c_void_p(buf.data_ptr()), ctypes.c_int(cta))
lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=5)
torch.cuda.synchronize()
l.osd_debug_attn_fwd_trace(ctypes.c_void_p(0), ctypes.c_int(0))
rec = buf.cpu().numpy().astype('uint64').reshape(2, 1024)
names = {0: 'wait_s', 1: 'got_s', 2: 'computed', 3: 'got_o', 4: 'arrived', 10: 'P_seen', 11: 'V_ready', 12: 'issued'}
ev = []
for slot in range(2):
    for r in rec[slot]:
        r = int(r)
        if r:
            ev.append((r & 0xffffffff, slot, (r >> 48) & 0xffff, (r >> 32) & 0xffff))
ev.sort()
t0 = ev[0][0]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (40, 43)
for t, slot, e, i in ev:
    if lo <= i <= hi:
        print(f'{t - t0:9d}  {"issuer" if slot == 0 else "warp2":7s} {names.get(e, e):9s} tile {i}')
ts = [t for t, s, e, i in ev if s == 1 and e == 1]
d = [b - a for a, b in zip(ts, ts[1:])]
print(f'steps {len(ts)}, mean period {sum(d) / len(d):.0f} clk, min {min(d)}, max {max(d)}')
for a, b, nm in ((0, 1, 'wait s_full'), (1, 2, 'compute'), (2, 3, 'wait o_ready'), (3, 4, 'store P + arrive')):
    ta = {i: t for t, s, e, i in ev if s == 1 and e == a}
    tb = {i: t for t, s, e, i in ev if s == 1 and e == b}
    dd = [tb[i] - ta[i] for i in ta if i in tb and i >= 4]
    print(f'warp2 {nm:18s}: mean {sum(dd) / len(dd):7.0f} clk')
for a, b, nm in ((10, 11, 'issuer wait V'), (11, 12, 'issuer PV + S issue')):
    ta = {i: t for t, s, e, i in ev if s == 0 and e == a}
    tb = {i: t for t, s, e, i in ev if s == 0 and e == b}
    dd = [tb[i] - ta[i] for i in ta if i in tb and i >= 4]
    print(f'{nm:24s}: mean {sum(dd) / len(dd):7.0f} clk')
