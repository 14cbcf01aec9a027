"""CUPTI kernel breakdown of ONE inference forward (torch.profiler CUDA activity; nsys is not in the image).
    python tools/forward_timeline.py [bf16|fp32] [B] [L]  ->  JSON on stdout: per kernel name launches and ms"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from oracle import denoiser_oracle as O  # seeded synthetic weights / inputs only
from osu_dreamer_b200.denoiser import DiffusionModel, default_args

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
m = DiffusionModel(6, 128, 32, default_args())
m.load_state_dict(O.make_state_dict(1234))
m = m.cuda().eval()
m.precision = prec
inp = {k: v.cuda() for k, v in O.make_inputs(B, L, seed=3).items()}
with torch.no_grad():
    for _ in range(3):
        m(inp['h'], inp['s'], inp['x0'])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(inp['h'], inp['s'], inp['x0'])
        torch.cuda.synchronize()
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, 't.json')
    prof.export_chrome_trace(path)
    trace = json.load(open(path))
ks = [(e['ts'], e['dur'], e['name']) for e in trace['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy') and 'dur' in e]
ks.sort()
agg = {}
for ts, dur, name in ks:
    short = name.split('(')[0].replace('void ', '').replace('osd::', '')[:70]
    d = agg.setdefault(short, [0, 0.0])
    d[0] += 1
    d[1] += dur
span = ks[-1][0] + ks[-1][1] - ks[0][0]
out = {'precision': prec, 'B': B, 'L': L, 'span_ms': span / 1e3, 'sum_kernel_ms': sum(d for _, d, _ in ks) / 1e3, 'launches': len(ks),
       'kernels': [{'kernel': k, 'launches': v[0], 'ms': round(v[1] / 1e3, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
