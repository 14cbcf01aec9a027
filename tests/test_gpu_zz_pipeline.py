"""GPU: the whole inference pipeline `LDM.sample` (osu_dreamer/models/inference/model.py:34-52) on the B200 path --
latent.audio_encoder -> style.sample -> diffusion.sample -> latent.decode -- against the CPU oracles composed the same way
on identical noise (fp32-grade denoiser, tolerance 5e-3), and `LDM.sample` itself against the manual composition under
the same seed (padding to the chunk size, noise draws in the reference's order, cropping)."""
import json
import os

import pytest
import torch

from oracle import denoiser_oracle as O
from oracle import neighbours_oracle as N

pytestmark = pytest.mark.gpu

HPARAMS = dict(emb_dim=6, style_dim=32, n_downs=3, stride=3,
               latent_args=dict(h_dim=128, ae_args=dict(n_layers=8, expand=4, radius=2), style_head_dim=64, style_heads=16),
               style_args=dict(label_features=128, h_dim=256, depth=8, expand=4),
               diffusion_args=dict(global_cond_dim=512, backbone_dim=512, u_head_dim=64,
                                   backbone_args=dict(depth=8, expand=4, head_dim=64, n_heads=16, radius=2)))


def _maxnorm(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def test_ldm_sample_matches_composed_oracles(golden_dir):
    from osu_dreamer_b200.ldm import LDM, pad_to_multiple
    spec = json.load(open(os.path.join(golden_dir, 'nb_spec.json')))
    lsd, ssd, dsd = N.seeded_state_dict(spec['latent'], 4321), N.seeded_state_dict(spec['style'], 8765), O.make_state_dict(1234)
    m = LDM(dict(HPARAMS))
    m.load_state_dict({**{'latent.' + k: v for k, v in lsd.items()}, **{'style.' + k: v for k, v in ssd.items()},
                       **{'diffusion.' + k: v for k, v in dsd.items()}}, strict=True)
    m = m.cuda().eval()
    m.diffusion.precision = 'fp32'  # fp32-grade denoiser: the comparison below is then an fp32-class one
    m.diffusion.graph_sampler = False  # eager launches (graph replay has its own test, tests/test_gpu_forward.py)
    g = torch.Generator().manual_seed(21)
    L, B, steps = 27 * 64 - 7, 2, 3
    audio = torch.randn(72, L, generator=g)
    labels = 10 * torch.rand(B, 5, generator=g)
    labels[1, 3] = -1.0
    s_init, x_init = torch.randn(B, 32, generator=g), torch.randn(B, 6, 64, generator=g)

    # ---- CPU oracles, fp64, composed like models/inference/model.py:45-52
    with torch.no_grad():
        ap = pad_to_multiple(audio, 27)
        skips_o, h_o = N.audio_encoder({k: v.double() for k, v in lsd.items()}, ap[None].double())
        s_o, _, _ = N.style_sample({k: v.double() for k, v in ssd.items()}, labels.double(), s_init.double(), 16)
        z_o, _, _ = O.sample({k: v.double() for k, v in dsd.items()}, h_o, s_o, x_init.double(), steps)
        chart_o, lab_o = N.decode({k: v.double() for k, v in lsd.items()}, z_o, s_o, skips_o)
        chart_o = chart_o[..., :L]

    # ---- the same composition on the B200 path with the same noise
    skips, h = m.latent.audio_encoder(pad_to_multiple(audio.cuda(), 27)[None])
    s = m.style.sample_from(labels.cuda(), s_init.cuda(), 16)
    z = m.diffusion.sample_from(h, s, x_init.cuda(), steps)
    chart, lab = m.latent.decode(z, s, skips=skips)
    chart = chart[..., :L]
    torch.cuda.synchronize()
    errs = dict(h=_maxnorm(h, h_o), s=_maxnorm(s, s_o), z=_maxnorm(z, z_o), chart=_maxnorm(chart, chart_o), labels=_maxnorm(lab, lab_o))
    print('pipeline', {k: f'{v:.2e}' for k, v in errs.items()})
    assert all(v < 5e-3 for v in errs.values())
    assert chart.shape == (B, 9, L) and float(chart[:, :7].min()) >= 0 and float(chart[:, :7].max()) <= 1

    # ---- LDM.sample: same seed -> same noise draws (style first, then the denoiser), padding and cropping inside
    torch.manual_seed(5)
    chart_a, lab_a = m.sample(audio.cuda(), labels.cuda(), steps)
    torch.manual_seed(5)
    s0 = torch.randn(B, 32, device='cuda')
    s1 = m.style.sample_from(labels.cuda(), s0, 16)
    x0 = torch.randn(B, 6, 64, device='cuda')
    z1 = m.diffusion.sample_from(h, s1, x0, steps)
    chart_b, lab_b = m.latent.decode(z1, s1, skips=skips)
    assert chart_a.shape == (B, 9, L) and torch.equal(chart_a, chart_b[..., :L]) and torch.equal(lab_a, lab_b)
