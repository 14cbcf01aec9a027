"""End-to-end `LDM.sample` latency on the B200 path for one song (the `predict` workload: scripts/predict.py:71-75),
split by stage: latent.audio_encoder, style.sample, diffusion.sample, latent.decode.  Synthetic weights and audio.
Usage: python tools/predict_latency.py [minutes=3] [n_diffs=4] [steps=8]   (6 ms per audio frame, 27 frames per latent)"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import denoiser_oracle as O
from oracle import neighbours_oracle as N
from osu_dreamer_b200.ldm import LDM, pad_to_multiple

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = json.load(open(os.path.join(root, 'tests', 'golden', 'nb_spec.json')))
hp = dict(emb_dim=6, style_dim=32, n_downs=3, stride=3,
          latent_args=dict(h_dim=128, ae_args=dict(n_layers=8, expand=4, radius=2), style_head_dim=64, style_heads=16),
          style_args=dict(label_features=128, h_dim=256, depth=8, expand=4),
          diffusion_args=dict(global_cond_dim=512, backbone_dim=512, u_head_dim=64,
                              backbone_args=dict(depth=8, expand=4, head_dim=64, n_heads=16, radius=2)))
m = LDM(dict(hp))
m.load_state_dict({**{'latent.' + k: v for k, v in N.seeded_state_dict(spec['latent'], 4321).items()},
                   **{'style.' + k: v for k, v in N.seeded_state_dict(spec['style'], 8765).items()},
                   **{'diffusion.' + k: v for k, v in O.make_state_dict(1234).items()}})
m = m.cuda().eval()
L = int(minutes * 60 / 0.006)
audio = torch.randn(72, L, device='cuda')
labels = 10 * torch.rand(B, 5, device='cuda')


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


ap = pad_to_multiple(audio, 27)[None]
t_enc, (skips, h) = timed(lambda: m.latent.audio_encoder(ap))
t_sty, s = timed(lambda: m.style.sample(labels))
t_dif, z = timed(lambda: m.diffusion.sample(h, s, steps))
t_dec, _ = timed(lambda: m.latent.decode(z, s, skips=skips))
t_all, (chart, lab) = timed(lambda: m.sample(audio, labels, steps))
row = {'song_minutes': minutes, 'audio_frames': L, 'latent_frames': h.shape[-1], 'difficulties': B, 'diffusion_steps': steps,
       'ms': {'audio_encoder': round(t_enc, 2), 'style_sample': round(t_sty, 2), 'diffusion_sample': round(t_dif, 2),
              'decode': round(t_dec, 2), 'LDM.sample': round(t_all, 2)},
       'finite': bool(torch.isfinite(chart).all())}
print(json.dumps(row))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(row, open('gpurun_out/predict_latency.json', 'w'), indent=1)
