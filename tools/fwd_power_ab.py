"""Sustained-loop timing of forward attention variants with the SM clock and board power sampled DURING each loop
(nvidia-smi, 100 ms): separates "fewer cycles" from "fewer joules" on a power-capped part.  -> gpurun_out/fwd_power_ab.json"""
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from osu_dreamer_b200 import lib


class Smi:
    def __init__(self):
        self.lines = []
        self.p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-lms', '100', '-i', '0'],
                                  stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for ln in self.p.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def window(self, t0, t1):
        v = [tuple(float(x) for x in ln.split(',')) for t, ln in self.lines if t0 + 0.3 < t < t1]
        if not v:
            return None
        return {'sm_mhz': sorted(a for a, _ in v)[len(v) // 2], 'power_w': sorted(b for _, b in v)[len(v) // 2], 'n': len(v)}


B, L = 16, 8192
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
q, k, v = (torch.randn(B, 16, L, 64, device='cuda').to(torch.bfloat16) for _ in range(3))
smi = Smi()
time.sleep(1.0)
res = []
variants = [int(a) for a in sys.argv[1:]] or [7, 15]


def run(name, fn, secs=2.5):
    fn()
    torch.cuda.synchronize()
    time.sleep(1.5)  # cool-down so every variant starts from a similar state
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    n = 0
    e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / n
    w = smi.window(t0, t1)
    d = {'name': name, 'ms': ms, 'tflops': 4.0 * B * 16 * L * L * 64 / ms / 1e9, **(w or {})}
    if w:
        d['mcycles'] = ms * 1e-3 * w['sm_mhz'] * 1e6 / 1e6
        d['joules_per_launch'] = w['power_w'] * ms * 1e-3
    res.append(d)
    print(json.dumps(d), flush=True)


for rep in range(2):
    for var in variants:
        run(f'variant {var}', lambda: lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=var))
    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]), torch.no_grad():
        run('cudnn sdpa fwd', lambda: F.scaled_dot_product_attention(q, k, v))
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'fwd_power_ab.json'), 'w'), indent=1)
smi.p.terminate()
