"""Sustained-loop timing of the attention backward with clock / power sampled during the loop (see fwd_power_ab.py), and
timing experiments that switch parts of the kernel off (OSD_FB_SKIP; results are then wrong, only the time / energy count).
Each setting in its own process.  -> gpurun_out/bwd_power_ab.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time, subprocess, threading
sys.path.insert(0, %r)
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel
from osu_dreamer_b200 import lib
lines = []
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '100', '-i', '0'], stdout=subprocess.PIPE, text=True)
def rd():
    for ln in p.stdout: lines.append((time.perf_counter(), ln.strip()))
threading.Thread(target=rd, daemon=True).start()
B, L = 16, 8192
which = sys.argv[1]
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
dy = torch.randn(B * L, 1024, device='cuda').to(torch.bfloat16)
y, lse = lib.attn_fwd(qkv, B, L)
if which == 'cudnn':
    q, k, v = (torch.randn(B, 16, L, 64, device='cuda').to(torch.bfloat16).requires_grad_(True) for _ in range(3))
    do = torch.randn(B, 16, L, 64, device='cuda').to(torch.bfloat16)
    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
        o = F.scaled_dot_product_attention(q, k, v)
    def fn():
        torch.autograd.grad(o, (q, k, v), do, retain_graph=True)
elif which == '2pass':
    fn = lambda: lib.attn_bwd(qkv, y, dy, lse, B, L)
else:
    fn = lambda: lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
fn(); torch.cuda.synchronize(); time.sleep(1.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); n = 0
e0.record()
while time.perf_counter() - t0 < 3.0:
    for _ in range(10): fn()
    n += 10; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
ms = e0.elapsed_time(e1) / n
v = [tuple(float(x) for x in ln.split(',')) for t, ln in lines if t0 + 0.4 < t < t1]
mhz = sorted(a for a, _ in v)[len(v) // 2]; w = sorted(b for _, b in v)[len(v) // 2]
print('RESULT ' + json.dumps({'ms': ms, 'tflops_algorithmic': 8.0 * B * 16 * L * L * 64 / ms / 1e9, 'sm_mhz': mhz, 'power_w': w, 'joules_per_call': w * ms * 1e-3, 'mcycles': ms * mhz * 1e-3}))
p.terminate()
''' % ROOT
res = []
SETS = (('fused', 'fused', {}), ('fused, 16 softmax warps', 'fused', {'OSD_FB_W16': '1'}), ('fused', 'fused', {}),
        ('fused, 16 softmax warps', 'fused', {'OSD_FB_W16': '1'}))
if len(sys.argv) > 1 and sys.argv[1] == 'all':
    SETS = (('fused', 'fused', {}), ('fused, no dQ reduce-add (timing only)', 'fused', {'OSD_FB_SKIP': '1'}), ('two-pass', '2pass', {}),
            ('cudnn sdpa bwd', 'cudnn', {}), ('fused', 'fused', {}))
for name, which, env in SETS:
    r = subprocess.run([sys.executable, '-c', CHILD, which], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    d = json.loads(line[0][7:]) if line else {'error': (r.stderr or r.stdout)[-400:]}
    d['name'] = name
    res.append(d)
    print(json.dumps(d), flush=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'bwd_power_ab.json'), 'w'), indent=1)
