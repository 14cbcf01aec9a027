"""A/B of the forward attention kernel's anti-phase start (csrc/attn_fwd_db.cu, g_db_sm_arrivals): OSD_FWD_STAGGER clocks x
library builds (OSD_LIB_PATH; e.g. the OSD_DB_EMU=1 build with a third of the exponentials on the FMA pipe).
Each setting runs in its own process (the settings are read once per process): value check + timing at B=16/32, L=8192."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
from osu_dreamer_b200 import lib
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
B, L = 2, 1000
g = torch.Generator().manual_seed(1)
qkv = torch.randn(B * L, 3072, generator=g).cuda().to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=7)
q, k, v = qkv.float().view(B, L, 3, 16, 64).permute(2, 0, 3, 1, 4)
s = (q @ k.transpose(-1, -2)) / 8
ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, 1024)
out = {'y_err': float((y.float() - ref).abs().max() / ref.abs().max()), 'lse_err': float((lse - torch.logsumexp(s, -1)).abs().max())}
for B, L in ((16, 8192), (32, 8192), (8, 2048)):
    qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
    ts = [timeit(lambda: lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=7)) for _ in range(3)]
    out[f'B{B}_L{L}_ms'] = [round(t, 4) for t in ts]
    out[f'B{B}_L{L}_tflops'] = round(4.0 * B * 16 * L * L * 64 / min(ts) / 1e9, 1)
print('RESULT ' + json.dumps(out))
''' % ROOT

res = []
libs = [None] + [p for p in sys.argv[1:]]
for libp in libs:
    for stag in (0, 250, 400, 550, 700, 900):
        env = dict(os.environ, OSD_FWD_STAGGER=str(stag))
        if libp:
            env['OSD_LIB_PATH'] = os.path.join(ROOT, libp)
        r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
        d = json.loads(line[0][7:]) if line else {'error': (r.stderr or r.stdout)[-400:]}
        d.update(stagger=stag, lib=libp or 'default')
        res.append(d)
        print(json.dumps(d), flush=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'fwd_stagger_ab.json'), 'w'), indent=1)
