"""64-step sampler at `predict`-like shapes (one song: l ~ 1-2 k latent frames, a few difficulties): eager launches
vs the captured CUDA graph.  Prints one JSON line per shape."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import denoiser_oracle as O
from osu_dreamer_b200.denoiser import DiffusionModel, default_args

m = DiffusionModel(6, 128, 32, default_args())
m.load_state_dict(O.make_state_dict(1234))
m = m.cuda().eval()
rows = []
for B, L in ((1, 1024), (4, 1024), (4, 2048), (8, 4096)):
    h = torch.randn(1, 128, L, device='cuda')
    s = torch.randn(B, 32, device='cuda')
    x0 = torch.randn(B, 6, L, device='cuda')
    row = {'B': B, 'L': L, 'num_steps': 64}
    for mode in (False, True):
        m.graph_sampler = mode
        for _ in range(3):
            m.sample_from(h, s, x0, 64)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            m.sample_from(h, s, x0, 64)
        torch.cuda.synchronize()
        row['graph_ms' if mode else 'eager_ms'] = (time.perf_counter() - t0) / n * 1e3
    row['speedup'] = row['eager_ms'] / row['graph_ms']
    row['latents_per_s_graph'] = B / row['graph_ms'] * 1e3
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/sampler_latency.json', 'w'), indent=1)
