"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the two neighbours of `diffusion.sample` inside `LDM.sample`
(osu_dreamer/models/inference/model.py:45-52):

  * the latent model's inference half: `audio_encoder` (SpecFeatures + UNetEncoder) before the sampler and `decode`
    (proj_emb -> UNetDecoder -> proj_out, label_predictor) after it;
  * the style model: `forward` and its sphere-tracing `sample`.

Functional, state-dict driven torch-CPU restatement; nothing in the product package may import this file.  Pinned
against outputs of the UNMODIFIED reference modules (`oracle/make_golden_neighbours.py` -> tests/golden/nb_*.npz, with
the parameter names / shapes in tests/golden/nb_spec.json); the reference ships no tests for these paths.  The CUDA
implementation of these rows is the next round's work: this file and its fixtures are the parity gate it will be held
to.  Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

A_DIM, X_DIM, NUM_LABELS = 72, 9, 5          # data/load_audio.py (A_DIM), data/beatmap/encode.py:29,50
HIT_SIGNALS, CURSOR_SIGNALS = list(range(7)), [7, 8]   # data/beatmap/encode.py (ONSET..CLAP | X, Y)
# models/latent/model.yml:88-101 and models/style/model.yml:71-75
LATENT_HP = dict(emb_dim=6, style_dim=32, n_downs=3, stride=3, h_dim=128, n_layers=8, expand=4, radius=2)
STYLE_HP = dict(style_dim=32, label_features=128, h_dim=256, depth=8, expand=4)


def seeded_state_dict(spec: Sequence[Tuple[str, Sequence[int]]], seed: int, dtype=torch.float32) -> SD:
    """Deterministic NON-DEGENERATE values for every entry of a state dict (the reference zero-initialises the FiLM
    layers, the mixer gates and several heads, which would make parity vacuous): tensors whose name ends in `gamma` or
    `.weight` with one dimension -> 1 + N(0, 0.1^2); other `weight` / `*_w` / `null_labels` -> N(0, 1/fan_in) with
    fan_in = prod(shape[1:]) (shape[-2] for the stacked cond_proj_w); `bias` / `*_b` -> N(0, 0.02^2); the Fourier
    feature buffers keep their reference distributions (common/fourier_features.py:11-12)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in spec:
        shape = tuple(shape)
        leaf = name.rsplit('.', 1)[-1]
        if leaf == 'gamma' or (leaf == 'weight' and len(shape) == 1):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith('rff.W'):
            t = torch.randn(shape, generator=g) * 32.0
        elif name.endswith('rff.b'):
            t = (torch.rand(shape, generator=g) * 2 - 1) * math.pi
        elif leaf == 'bias' or leaf.endswith('_b'):
            t = torch.randn(shape, generator=g) * 0.02
        else:
            fan_in = shape[-2] if leaf == 'cond_proj_w' else (math.prod(shape[1:]) if len(shape) > 1 else shape[0])
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        if leaf == 'bias' and name.endswith('u_out.bias'):
            t = t - 0.4328  # the reference's operating point (models/style/model.py:57)
        sd[name] = t.to(dtype)
    return sd


# ---------------------------------------------------------------------------------------------- shared ops
def rms_norm(x: Tensor, gamma: Tensor | None = None) -> Tensor:
    """common/rms_norm.py:7-16 -- over dim 1, eps 1e-6, gamma broadcast over the trailing dims; the statistic is
    formed in fp32 whatever the input dtype (`x.float()`), so the reference's fp64 runs carry fp32 norms too."""
    xf = x.float()
    n = (xf * xf.pow(2).mean(dim=1, keepdim=True).add(1e-6).rsqrt()).to(x.dtype)
    if gamma is not None:
        n = n * gamma.view(-1, *([1] * (x.ndim - 2)))
    return n


def swiglu(sd: SD, p: str, x: Tensor, radius: int) -> Tensor:
    """common/swiglu.py:27-32 (depthwise conv k = 1 + 2 radius, 1x1 to 2 h, v * silu(g), RMSNorm(h, no affine), 1x1)."""
    h = F.conv1d(x, sd[p + 'proj_vg.0.weight'], sd[p + 'proj_vg.0.bias'], padding=radius, groups=x.shape[1])
    v, g = F.conv1d(h, sd[p + 'proj_vg.1.weight'], sd[p + 'proj_vg.1.bias']).chunk(2, dim=1)
    return F.conv1d(rms_norm(v * F.silu(g)), sd[p + 'proj_o.weight'], sd[p + 'proj_o.bias'])


def layer(sd: SD, p: str, x: Tensor, cond: Tensor | None, hp=LATENT_HP) -> Tensor:
    """models/latent/unet.py:22-55: n_layers x [x += RMSNorm_gain(SwiGLU(norm(x)(1+scale)+shift)) (1+gate)], out_norm."""
    for j in range(hp['n_layers']):
        if cond is not None:
            scale, shift, gate = F.linear(cond, sd[f'{p}films.{j}.weight'], sd[f'{p}films.{j}.bias'])[:, :, None].chunk(3, dim=1)
        else:
            scale = shift = gate = 0
        h = rms_norm(x, sd[f'{p}norms.{j}.gamma']) * (1 + scale) + shift
        h = rms_norm(swiglu(sd, f'{p}blocks.{j}.0.', h, hp['radius']), sd[f'{p}blocks.{j}.1.gamma'])
        x = x + h * (1 + gate)
    return rms_norm(x, sd[p + 'out_norm.gamma'])


# ---------------------------------------------------------------------------------------------- latent model
def spec_features(sd: SD, p: str, x: Tensor) -> Tensor:
    """models/latent/spec_features.py:18-30: [B,F,L] -> [B,128,L]."""
    h = x.unsqueeze(1)
    h = F.silu(rms_norm(F.conv2d(h, sd[p + 'net.1.weight'], sd[p + 'net.1.bias'], stride=(6, 1), padding=(1, 1)), sd[p + 'net.2.gamma']))
    h = F.silu(rms_norm(F.conv2d(h, sd[p + 'net.4.weight'], sd[p + 'net.4.bias'], stride=(4, 1), padding=(1, 1)), sd[p + 'net.5.gamma']))
    h = h.flatten(1, 2)  # 'b c a l -> b (c a) l'
    return F.silu(rms_norm(F.conv1d(h, sd[p + 'net.8.weight'], sd[p + 'net.8.bias']), sd[p + 'net.9.gamma']))


def unet_encoder(sd: SD, p: str, x: Tensor, hp=LATENT_HP) -> Tuple[List[Tensor], Tensor]:
    """models/latent/unet.py:57-76: per level layer -> (skip = x) -> depthwise conv k=3 + AvgPool(stride)."""
    skips = []
    for i in range(hp['n_downs']):
        x = layer(sd, f'{p}layers.{i}.', x, None, hp)
        skips.append(x)
        x = F.conv1d(x, sd[f'{p}downs.{i}.0.weight'], sd[f'{p}downs.{i}.0.bias'], padding=hp['stride'] // 2, groups=x.shape[1])
        x = F.avg_pool1d(x, hp['stride'])
    return skips, x


def unet_decoder(sd: SD, p: str, skips: List[Tensor], x: Tensor, cond: Tensor, hp=LATENT_HP) -> Tensor:
    """models/latent/unet.py:78-101 (+ mixer :115-126): nearest upsample x stride, depthwise conv k=3,
    x += RMSNorm(conv1x1(skip)) * conv1x1(x), conditional layer."""
    skips = list(skips)
    for i in range(hp['n_downs']):
        x = F.interpolate(x, scale_factor=hp['stride'], mode='nearest')
        x = F.conv1d(x, sd[f'{p}ups.{i}.1.weight'], sd[f'{p}ups.{i}.1.bias'], padding=hp['stride'] // 2, groups=x.shape[1])
        skip = skips.pop().expand(x.size(0), -1, -1)
        pr = rms_norm(F.conv1d(skip, sd[f'{p}mixers.{i}.proj.0.weight'], sd[f'{p}mixers.{i}.proj.0.bias']),
                      sd[f'{p}mixers.{i}.proj.1.gamma'])
        x = x + pr * F.conv1d(x, sd[f'{p}mixers.{i}.gate.weight'], sd[f'{p}mixers.{i}.gate.bias'])
        x = layer(sd, f'{p}layers.{i}.', x, cond, hp)
    return x


def audio_encoder(sd: SD, audio: Tensor, hp=LATENT_HP) -> Tuple[List[Tensor], Tensor]:
    """LatentModel.audio_encoder (models/latent/model.py:53): audio [B,72,L] -> (skips, h [B,128,L/27])."""
    return unet_encoder(sd, 'audio_encoder.1.', spec_features(sd, 'audio_encoder.0.', audio), hp)


def decode(sd: SD, z: Tensor, s: Tensor, skips: List[Tensor], hp=LATENT_HP) -> Tuple[Tensor, Tensor]:
    """LatentModel.decode (models/latent/model.py:103-133): chart [B,9,L] (sigmoid on the hit signals), labels [B,5]."""
    x = F.conv1d(z, sd['proj_emb.weight'], sd['proj_emb.bias'])
    logits = F.conv1d(unet_decoder(sd, 'decoder.', skips, x, s, hp), sd['proj_out.weight'], sd['proj_out.bias'])
    chart = torch.cat([logits[:, HIT_SIGNALS].sigmoid(), logits[:, CURSOR_SIGNALS]], dim=1)
    lab = F.linear(F.silu(F.linear(s, sd['label_predictor.0.weight'], sd['label_predictor.0.bias'])),
                   sd['label_predictor.2.weight'], sd['label_predictor.2.bias']).clamp(0, 10)
    return chart, lab


# ---------------------------------------------------------------------------------------------- style model
def style_constants(style_dim: int = 32) -> Tuple[float, float]:
    """c0, u_scale -- models/style/model.py:34-40."""
    d0_sq = 2.0 * style_dim
    t99 = torch.tensor(2.3263478740408408).sigmoid().item()
    return (1 - t99) ** 2 * d0_sq, math.sqrt(d0_sq)


def style_conditioning(sd: SD, labels: Tensor, hp=STYLE_HP) -> Tensor:
    """models/style/model.py:72-79 + common/fourier_features.py:15-16; labels < 0 select the learned null embedding."""
    lab = labels[:, :, None]
    ff = (2 / hp['label_features']) ** 0.5 * torch.cos((lab / 10) @ sd['rff.W'].T + sd['rff.b'])
    h = torch.einsum('bnf,nfh->bnh', ff, sd['cond_proj_w']) + sd['cond_proj_b']
    return torch.where(lab < 0, sd['null_labels'][None], h).sum(dim=1)


def style_forward(sd: SD, st: Tensor, labels: Tensor, hp=STYLE_HP) -> Tuple[Tensor, Tensor]:
    """models/style/model.py:81-99 -> (u [B], v [B,S])."""
    c = style_conditioning(sd, labels, hp)
    x = F.linear(st, sd['proj_in.weight'], sd['proj_in.bias'])
    for i in range(hp['depth']):
        scale, shift, gate = F.linear(c, sd[f'films.{i}.weight'], sd[f'films.{i}.bias']).chunk(3, dim=1)
        h = rms_norm(x) * (1 + scale) + shift
        h = F.linear(F.silu(F.linear(h, sd[f'blocks.{i}.0.weight'], sd[f'blocks.{i}.0.bias'])),
                     sd[f'blocks.{i}.3.weight'], sd[f'blocks.{i}.3.bias'])
        x = x + rms_norm(h) * gate
    eps = torch.finfo(x.dtype).eps  # nn.RMSNorm(eps=None)
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * sd['proj_out.0.weight']
    v = F.linear(xn, sd['proj_out.1.weight'], sd['proj_out.1.bias'])
    u_scale = style_constants(hp['style_dim'])[1]
    u = u_scale * F.softplus(F.linear(rms_norm(x), sd['u_out.weight'], sd['u_out.bias'])).squeeze(-1)
    return u, v


def style_sample(sd: SD, labels: Tensor, s_init: Tensor, num_steps: int = 16, hp=STYLE_HP):
    """models/style/model.py:101-119 with the initial noise passed in -> (s [B,S], u0, eta)."""
    c0 = style_constants(hp['style_dim'])[0]
    s = s_init.clone()
    u0 = style_forward(sd, s, labels, hp)[0].mean().item()
    eta = 1.0 - (math.sqrt(c0) / max(u0, math.sqrt(c0) + 1e-6)) ** (1.0 / num_steps)
    for _ in range(num_steps):
        u, v = style_forward(sd, s, labels, hp)
        s = s - eta * u[:, None] * v
    return s, u0, eta


# ---------------------------------------------------------------------------------------------- style trainer loss
def style_trainer_loss(sd: SD, s1: Tensor, labels: Tensor, s0: Tensor, t: Tensor, drop: Tensor,
                       osl_weight: float = 1.0, del_weight: float = 30.0, hp=STYLE_HP):
    """StyleTrainer.forward (models/style/train.py:48-91) with its random draws passed in: t [B] (the stratified
    logit-normal times), s0 [B,S] (noise), drop [B,5] bool (labels replaced by -1) -> (loss, dict(osl, del, u_mape))."""
    c0 = style_constants(hp['style_dim'])[0]
    st = torch.lerp(s0, s1, t[:, None])
    masked = torch.where(drop, torch.full_like(labels, -1.0), labels)
    u_pred, v_pred = style_forward(sd, st, masked, hp)
    d_sq = (st - s1).square().sum(1)
    u_target = (d_sq + c0).sqrt()
    denoised = st - u_pred[:, None] * v_pred
    osl = ((denoised - s1).square().sum(1) / (d_sq + c0)).mean()
    v_target = (st - s1) / u_target[:, None]
    del_ = (v_pred - v_target).square().sum(1).mean()
    loss = osl_weight * osl + del_weight * del_
    u_err = ((u_pred - u_target) / u_target).abs().mean()
    return loss, dict(osl=osl, del_=del_, u_mape=u_err)
