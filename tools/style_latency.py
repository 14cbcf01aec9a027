"""Style sampler (16 steps = 17 forwards) latency on the B200 path vs the CPU oracle port of the reference arithmetic."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import neighbours_oracle as N
from osu_dreamer_b200.style import StyleModel, StyleModelArgs

spec = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'nb_spec.json')))['style']
sd = N.seeded_state_dict(spec, 8765)
m = StyleModel(32, StyleModelArgs(128, 256, 8, 4))
m.load_state_dict(sd)
m = m.cuda().eval()
rows = []
for B in (1, 4, 32, 256):
    labels = 10 * torch.rand(B, 5)
    lc = labels.cuda()
    for _ in range(3):
        m.sample(lc, 16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        m.sample(lc, 16)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / n * 1e3
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    s0 = torch.randn(B, 32)
    with torch.no_grad():
        N.style_sample(sd, labels, s0, 16)
        t0 = time.perf_counter()
        for _ in range(3):
            N.style_sample(sd, labels, s0, 16)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    row = {'B': B, 'num_steps': 16, 'b200_ms': round(gpu_ms, 3), 'cpu_oracle_ms': round(cpu_ms, 2), 'codes_per_s_b200': round(B / gpu_ms * 1e3, 1)}
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/style_latency.json', 'w'), indent=1)
