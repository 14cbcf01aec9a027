"""Where `latent.audio_encoder` / `latent.decode` spend their time for one 4-minute song, 4 difficulties: wall clock, summed
kernel time (CUPTI through torch.profiler) and the per-kernel totals.  Usage: python tools/latent_breakdown.py [tc|fp32]"""
import collections
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from oracle import neighbours_oracle as N
from osu_dreamer_b200.latent import LatentModel

impl = sys.argv[1] if len(sys.argv) > 1 else 'tc'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = json.load(open(os.path.join(root, 'tests', 'golden', 'nb_spec.json')))
m = LatentModel(6, 32, 3, 3, dict(h_dim=128, ae_args=dict(n_layers=8, expand=4, radius=2), style_head_dim=64, style_heads=16))
m.load_state_dict(N.seeded_state_dict(spec['latent'], 4321))
m = m.cuda().eval()
m.block_impl = impl
L = 40014  # 4 minutes at 6 ms per frame, padded to a multiple of 27
audio = torch.randn(1, 72, L, device='cuda')
z, s = torch.randn(4, 6, L // 27, device='cuda'), torch.randn(4, 32, device='cuda')
out = {'impl': impl}
for name, fn in (('audio_encoder', lambda: m.audio_encoder(audio)), ('decode', None)):
    if name == 'decode':
        skips, _ = m.audio_encoder(audio)
        fn = lambda: m.decode(z, s, skips=skips)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 3 * 1e3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    agg = collections.Counter()
    cnt = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name.split('(')[0].replace('void ', '').replace('osd::', '')[:60]
            agg[k] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
            cnt[k] += 1
    tot = sum(agg.values()) / 1e3
    out[name] = {'wall_ms': round(wall, 2), 'kernel_ms': round(tot, 2), 'launches': sum(cnt.values()),
                 'kernels': {k: [round(v / 1e3, 3), cnt[k]] for k, v in agg.most_common(12)}}
print(json.dumps(out, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open(f'gpurun_out/latent_breakdown_{impl}.json', 'w'), indent=1)
