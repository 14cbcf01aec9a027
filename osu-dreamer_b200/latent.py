"""Mirror of the reference's `LatentModel` (osu_dreamer/models/latent/model.py:39-133) for INFERENCE on the B200 path:
the two calls `LDM.sample` makes around `diffusion.sample` (models/inference/model.py:47,51) --
`latent.audio_encoder(audio)` -> (skips, h) and `latent.decode(z, s, skips=skips)` -> (chart, labels) -- with the
reference's full parameter tree (same names, shapes and order: the `latent.*` part of an inference artifact loads
strictly, including the chart encoder / style head / temporal head that only training uses).

The arithmetic runs in csrc/latent.cu through the C ABI (`osd_lat_*`, fp32, channels-first).  Training the latent model
(`fit-latent`) and `encode_chart` are not on this path.  No CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import sqrt

import torch
from torch import Tensor, nn

from . import lib

A_DIM, X_DIM, NUM_LABELS = 72, 9, 5   # data/load_audio.py, data/beatmap/encode.py:29,50
N_HIT = 7                             # HitSignals = ONSET..CLAP, then CursorSignals X, Y (encode.py)


@dataclass
class LayerArgs:  # models/latent/unet.py:9-13
    n_layers: int
    expand: int
    radius: int


@dataclass
class LatentModelArgs:  # models/latent/model.py:15-21
    h_dim: int
    ae_args: LayerArgs
    style_head_dim: int
    style_heads: int


class _Bag(nn.Module):
    """nameable container of parameters / sub-bags"""


def _put(root: nn.Module, name: str, shape, init, bag_types=None):
    """register parameter `name` (dotted) under nested bags; init: 'lin' U(+-1/sqrt(fan_in)), 'zero', or a float constant"""
    parts = name.split('.')
    m = root
    for d, p in enumerate(parts[:-1]):
        if not hasattr(m, p):
            setattr(m, p, (bag_types or {}).get(p, _Bag)() if d == 0 else _Bag())
        m = getattr(m, p)
    t = torch.empty(*shape)
    if init == 'lin':
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        nn.init.uniform_(t, -1 / sqrt(max(fan_in, 1)), 1 / sqrt(max(fan_in, 1)))
    elif init == 'zero':
        t.zero_()
    else:
        t.fill_(float(init))
    setattr(m, parts[-1], nn.Parameter(t))


def _conv(spec, p, cout, cin, k=None, init='lin'):
    spec.append((p + '.weight', (cout, cin) if k is None else (cout, cin, *k), init))
    spec.append((p + '.bias', (cout,), init))


def _layer_spec(spec, p, dim, cond_dim, a: LayerArgs):
    """unet.py:22-38 in registration order: norms, blocks, out_norm, films"""
    h = int(dim * a.expand * 2 / 3)
    for j in range(a.n_layers):
        spec.append((f'{p}.norms.{j}.gamma', (dim,), 1.0))
    for j in range(a.n_layers):
        b = f'{p}.blocks.{j}.0'
        if a.radius > 0:
            _conv(spec, b + '.proj_vg.0', dim, 1, (1 + 2 * a.radius,))
        _conv(spec, b + '.proj_vg.1', 2 * h, dim, (1,))
        _conv(spec, b + '.proj_o', dim, h, (1,))
        spec.append((f'{p}.blocks.{j}.1.gamma', (dim,), 1e-3))
    spec.append((f'{p}.out_norm.gamma', (dim,), 1.0))
    if cond_dim > 0:
        for j in range(a.n_layers):
            _conv(spec, f'{p}.films.{j}', 3 * dim, cond_dim, None, 'zero')


def _encoder_spec(spec, p, dim, n_downs, stride, a):
    for i in range(n_downs):
        _conv(spec, f'{p}.downs.{i}.0', dim, 1, (1 + 2 * (stride // 2),))
    for i in range(n_downs):
        _layer_spec(spec, f'{p}.layers.{i}', dim, 0, a)


def parameter_table(emb_dim, style_dim, n_downs, stride, args: LatentModelArgs):
    """(name, shape, init) of the whole LatentModel in the reference's state-dict order (model.py:53-77)"""
    D, a = args.h_dim, args.ae_args
    spec = []
    _conv(spec, 'chart_encoder.0', D, X_DIM, (1,))
    _encoder_spec(spec, 'chart_encoder.1', D, n_downs, stride, a)
    _conv(spec, 'audio_encoder.0.net.1', 8, 1, (8, 3))
    spec.append(('audio_encoder.0.net.2.gamma', (8,), 1.0))
    _conv(spec, 'audio_encoder.0.net.4', 32, 8, (6, 3))
    spec.append(('audio_encoder.0.net.5.gamma', (32,), 1.0))
    _conv(spec, 'audio_encoder.0.net.8', D, 32 * (A_DIM // 24), (1,))
    spec.append(('audio_encoder.0.net.9.gamma', (D,), 1.0))
    _encoder_spec(spec, 'audio_encoder.1', D, n_downs, stride, a)
    _layer_spec(spec, 'style_head.0', D, 0, a)
    hd = args.style_head_dim * args.style_heads
    _conv(spec, 'style_head.1.scores', args.style_heads, D, (1,))
    _conv(spec, 'style_head.1.values', hd, D, (1,))
    _conv(spec, 'style_head.1.proj_out', style_dim, hd)
    _layer_spec(spec, 'temporal_layer', D, style_dim, a)
    _conv(spec, 'temporal_head.0', emb_dim, D, (1,))
    _conv(spec, 'proj_emb', D, emb_dim, (1,))
    for i in range(n_downs):
        _conv(spec, f'decoder.ups.{i}.1', D, 1, (1 + 2 * (stride // 2),))
    for i in range(n_downs):
        _layer_spec(spec, f'decoder.layers.{i}', D, style_dim, a)
    for i in range(n_downs):
        _conv(spec, f'decoder.mixers.{i}.proj.0', D, D, (1,))
        spec.append((f'decoder.mixers.{i}.proj.1.gamma', (D,), 1.0))
        _conv(spec, f'decoder.mixers.{i}.gate', D, D, (1,), 'zero')
    _conv(spec, 'proj_out', X_DIM, D, (1,))
    _conv(spec, 'label_predictor.0', D, style_dim)
    _conv(spec, 'label_predictor.2', NUM_LABELS, D)
    return spec


class _AudioEncoder(_Bag):
    """holds audio_encoder.* and is callable like the reference's nn.Sequential(SpecFeatures, UNetEncoder)"""

    def forward(self, audio: Tensor):
        return self._owner[0]._audio_encoder(audio)


class LatentModel(nn.Module):
    def __init__(self, emb_dim: int, style_dim: int, n_downs: int, stride: int, args: LatentModelArgs):
        super().__init__()
        if isinstance(args, dict):
            args = dict(args)
            args['ae_args'] = LayerArgs(**args['ae_args']) if isinstance(args['ae_args'], dict) else args['ae_args']
            args = LatentModelArgs(**args)
        a = args.ae_args
        if (emb_dim, style_dim, n_downs, stride, args.h_dim, a.n_layers, a.expand, a.radius) != (6, 32, 3, 3, 128, 8, 4, 2):
            raise lib.OsdError('libosd_b200 is compiled for the latent model of models/latent/model.yml:84-101 '
                               '(emb 6, style 32, 3 downs of stride 3, h_dim 128, 8 layers, expand 4, radius 2)')
        self.emb_dim, self.style_dim, self.a_dim = emb_dim, style_dim, args.h_dim
        self.n_downs, self.stride, self.n_layers = n_downs, stride, a.n_layers
        self.chunk_size = stride ** n_downs
        self.block_impl = 'tc'  # 'fp32': the exact-fp32 CUDA-core block kernel (csrc/latent.cu)
        for name, shape, init in parameter_table(emb_dim, style_dim, n_downs, stride, args):
            _put(self, name, shape, init, {'audio_encoder': _AudioEncoder})
        object.__setattr__(self.audio_encoder, '_owner', (self,))  # not a submodule: no reference cycle in the module tree

    # ------------------------------------------------------------------ helpers
    def _sd(self):
        sd = {k: v for k, v in self.named_parameters()}
        if any(not v.is_cuda for v in sd.values()):
            raise lib.OsdError('LatentModel parameters must be on a CUDA device: libosd_b200 has no CPU path')
        return {k: v.detach() for k, v in sd.items()}

    def _packed(self, key: str, w1: Tensor, b1: Tensor, w2: Tensor) -> Tensor:
        """the block's GEMM weights as (hi | lo) tf32 splits, re-packed when a parameter was replaced or written in place"""
        cache = self.__dict__.setdefault('_tc_cache', {})
        tag = tuple((t.data_ptr(), t._version) for t in (w1, b1, w2))
        hit = cache.get(key)
        if hit is None or hit[0] != tag:
            hit = cache[key] = (tag, lib.lat_tc_pack(w1, b1, w2))
        return hit[1]

    def _layer(self, sd, p: str, x: Tensor, cond: Tensor | None) -> Tensor:
        """unet.py:40-55.  `block_impl`: 'tc' (default) = the two 1x1 convolutions as 3xTF32 tcgen05 GEMMs (split operands,
        ~1e-6 of fp32), 'fp32' = the exact-fp32 CUDA-core kernel."""
        if self.block_impl not in ('tc', 'fp32'):
            raise lib.OsdError(f"LatentModel.block_impl must be 'tc' or 'fp32', not {self.block_impl!r}")
        ws = lib.lat_tc_workspace(x.shape[0], x.shape[2], x.device) if self.block_impl == 'tc' else None
        for j in range(self.n_layers):
            film = lib.lat_conv1x1(cond, sd[f'{p}.films.{j}.weight'], sd[f'{p}.films.{j}.bias']) if cond is not None else None
            b = f'{p}.blocks.{j}.0'
            w8 = [sd[f'{p}.norms.{j}.gamma'], sd[b + '.proj_vg.0.weight'], sd[b + '.proj_vg.0.bias'], sd[b + '.proj_vg.1.weight'],
                  sd[b + '.proj_vg.1.bias'], sd[b + '.proj_o.weight'], sd[b + '.proj_o.bias'], sd[f'{p}.blocks.{j}.1.gamma']]
            if self.block_impl == 'tc':
                x = lib.lat_block_tc(x, w8, self._packed(f'{p}.{j}', w8[3], w8[4], w8[5]), film, ws)
            else:
                x = lib.lat_block(x, w8, film)
        return lib.lat_rmsnorm(x, sd[f'{p}.out_norm.gamma'])

    @torch.no_grad()
    def _audio_encoder(self, audio: Tensor):
        """SpecFeatures (spec_features.py:18-30) + UNetEncoder (unet.py:68-76): audio [B,72,L] -> (skips, h [B,128,L/27])"""
        sd = self._sd()
        p = 'audio_encoder.0.net'
        x = audio.float().contiguous()
        if x.dim() != 3 or x.shape[1] != A_DIM or x.shape[2] % self.chunk_size:
            raise lib.OsdError(f'audio must be [B, {A_DIM}, L] with L a multiple of {self.chunk_size}')
        h = lib.lat_conv2d(x.unsqueeze(1), sd[p + '.1.weight'], sd[p + '.1.bias'], 6)
        h = lib.lat_rmsnorm(h, sd[p + '.2.gamma'], silu=True)
        h = lib.lat_conv2d(h, sd[p + '.4.weight'], sd[p + '.4.bias'], 4)
        h = lib.lat_rmsnorm(h, sd[p + '.5.gamma'], silu=True)
        h = h.flatten(1, 2)  # 'b c a l -> b (c a) l' (contiguous view)
        h = lib.lat_rmsnorm(lib.lat_conv1x1(h, sd[p + '.8.weight'], sd[p + '.8.bias']), sd[p + '.9.gamma'], silu=True)
        skips = []
        for i in range(self.n_downs):
            h = self._layer(sd, f'audio_encoder.1.layers.{i}', h, None)
            skips.append(h)
            h = lib.lat_down3(h, sd[f'audio_encoder.1.downs.{i}.0.weight'], sd[f'audio_encoder.1.downs.{i}.0.bias'])
        return skips, h

    @torch.no_grad()
    def decode(self, z: Tensor, s: Tensor, *, audio: Tensor | None = None, skips=None):
        """model.py:118-133 -> (chart [B,9,L] with sigmoid on the hit signals, labels [B,5] clamped to [0,10])"""
        if skips is None:
            skips, _ = self._audio_encoder(audio)
        sd = self._sd()
        s = s.float().contiguous()
        x = lib.lat_conv1x1(z.float().contiguous(), sd['proj_emb.weight'], sd['proj_emb.bias'])
        skips = list(skips)
        for i in range(self.n_downs):
            x = lib.lat_up3(x, sd[f'decoder.ups.{i}.1.weight'], sd[f'decoder.ups.{i}.1.bias'])
            skip = skips.pop().float().contiguous()
            m = f'decoder.mixers.{i}'
            pr = lib.lat_rmsnorm(lib.lat_conv1x1(skip, sd[m + '.proj.0.weight'], sd[m + '.proj.0.bias']), sd[m + '.proj.1.gamma'])
            x = lib.lat_mix(x, pr, lib.lat_conv1x1(x, sd[m + '.gate.weight'], sd[m + '.gate.bias']))
            x = self._layer(sd, f'decoder.layers.{i}', x, s)
        chart = lib.lat_conv1x1(x, sd['proj_out.weight'], sd['proj_out.bias'], act=2, act_channels=N_HIT)
        lab = lib.lat_conv1x1(lib.lat_conv1x1(s, sd['label_predictor.0.weight'], sd['label_predictor.0.bias'], act=1),
                              sd['label_predictor.2.weight'], sd['label_predictor.2.bias']).clamp(0, 10)
        return chart, lab

    def forward(self, *a, **k):
        raise lib.OsdError('the B200 latent path is inference-only (audio_encoder / decode); fit-latent stays on the reference')
