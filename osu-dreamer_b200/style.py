"""Mirror of the reference's `StyleModel` (osu_dreamer/models/style/model.py:19-119) on the B200 path: same constructor,
attributes (`style_dim`, `c0`, `u_scale`), parameter / buffer names and shapes (so the `style.*` part of an inference
artifact loads strictly), `forward(st, labels) -> (u, v)` and the sphere-tracing `sample(labels, num_steps=16)` that
`LDM.sample` calls right before `diffusion.sample` (models/inference/model.py:48-49).

Inference runs in csrc/style.cu through the C ABI (`osd_style_forward`, `osd_style_sample`): one CTA per sample, all
sampler steps inside one launch.  Under autograd (`fit-style`, style_trainer.py) `forward` is one autograd Function
whose forward / backward are `osd_style_train_forward` / `osd_style_backward` (csrc/style_train.cu).  No CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import sqrt

import torch
from torch import Tensor, nn

from . import lib

NUM_LABELS = 5  # osu_dreamer/data/beatmap/encode.py:50


@dataclass
class StyleModelArgs:  # models/style/model.py:19-25
    label_features: int
    h_dim: int
    depth: int
    expand: int
    dropout: float = 0.


class _FourierFeatures(nn.Module):  # common/fourier_features.py:7-13 (buffers only; evaluated inside the kernel)
    def __init__(self, dim: int, features: int, n_bins: int = 16):
        super().__init__()
        self.register_buffer('W', torch.randn(features, dim) * float(n_bins))
        self.register_buffer('b', torch.empty(features).uniform_(-torch.pi, torch.pi))


class _StyleFn(torch.autograd.Function):
    """StyleModel.forward under autograd: forward = osd_style_train_forward (activations kept in a workspace), backward =
    osd_style_backward (parameter gradients; accumulated straight into the trainer's flat buffer when it lent one)."""

    @staticmethod
    def forward(ctx, model, st, labels, *params):
        parr, keep = model._parr()
        ws = lib.style_train_workspace(st.shape[0], st.device)
        u, v = lib.style_train_forward(parr, st, labels, ws)
        ctx.model, ctx.keep = model, (parr, keep, st, labels, ws)
        return u, v

    @staticmethod
    def backward(ctx, du, dv):
        model = ctx.model
        parr, keep, st, labels, ws = ctx.keep
        B, dev = st.shape[0], st.device
        params = list(model.parameters())
        targets = getattr(model, '_grad_targets', None)
        direct = targets is not None and len(targets) == len(params) and targets[0].device == dev
        if direct:
            grads = targets
        else:
            sizes = [(p.numel() + 63) // 64 * 64 for p in params]
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            grads, off = [], 0
            for p, n in zip(params, sizes):
                grads.append(flat[off:off + p.numel()].view(p.shape))
                off += n
        # state-dict order = parameters with the two Fourier-feature buffers after null_labels (model.py:42-46)
        slots = grads[:3] + [None, None] + grads[3:]
        du = torch.zeros(B, device=dev) if du is None else du.float().contiguous()
        dv = torch.zeros(B, model.style_dim, device=dev) if dv is None else dv.float().contiguous()
        lib.style_backward(parr, st, labels, du, dv, lib.style_grad_array(slots), ws)
        return (None, None, None, *([None] * len(params) if direct else grads))


class StyleModel(nn.Module):
    def __init__(self, style_dim: int, args: StyleModelArgs):
        super().__init__()
        if isinstance(args, dict):
            args = StyleModelArgs(**args)
        if (style_dim, args.label_features, args.h_dim, args.depth, args.expand) != (32, 128, 256, 8, 4):
            raise lib.OsdError('libosd_b200 is compiled for the style model of models/style/model.yml:68-75 '
                               '(style_dim 32, label_features 128, h_dim 256, depth 8, expand 4)')
        self.style_dim = style_dim
        d0_sq = 2. * style_dim
        t99 = torch.tensor(2.3263478740408408).sigmoid().item()
        self.c0 = (1 - t99) ** 2 * d0_sq
        self.u_scale = sqrt(d0_sq)
        H = args.h_dim
        # same names / shapes / state-dict order as the reference (model.py:42-70)
        self.cond_proj_w = nn.Parameter(torch.empty(NUM_LABELS, args.label_features, H))
        self.cond_proj_b = nn.Parameter(torch.zeros(NUM_LABELS, H))
        for w in self.cond_proj_w:
            nn.init.xavier_uniform_(w)
        self.null_labels = nn.Parameter(torch.randn(NUM_LABELS, H) * H ** -.5)
        self.rff = _FourierFeatures(1, args.label_features, n_bins=32)
        self.proj_in = nn.Linear(style_dim, H)
        self.proj_out = nn.Sequential(nn.RMSNorm(H), nn.Linear(H, style_dim))
        nn.init.zeros_(self.proj_out[1].weight)
        nn.init.zeros_(self.proj_out[1].bias)
        self.u_out = nn.Linear(H, 1)
        nn.init.zeros_(self.u_out.weight)
        nn.init.constant_(self.u_out.bias, -0.4328)
        self.films = nn.ModuleList([nn.Linear(H, 3 * H) for _ in range(args.depth)])
        for f in self.films:
            nn.init.zeros_(f.weight)
            nn.init.zeros_(f.bias)
        self.blocks = nn.ModuleList([
            nn.Sequential(nn.Linear(H, args.expand * H), nn.SiLU(), nn.Dropout(args.dropout), nn.Linear(args.expand * H, H))
            for _ in range(args.depth)])

    def _tensors(self):
        sd = self.state_dict()
        return [sd[k] for k in sd]  # state-dict order == the library's parameter order (checked by the tests)

    def _parr(self):
        ts = self._tensors()
        if any(not t.is_cuda for t in ts):
            raise lib.OsdError('StyleModel tensors must be on a CUDA device: libosd_b200 has no CPU path')
        # the converted tensors (copies when a parameter is not already contiguous fp32, e.g. after .half()) must outlive
        # the enqueued kernels: the caller keeps the returned list until its library call has been issued
        conv = [t.detach().float().contiguous() for t in ts]
        return lib.style_param_array(conv), conv

    def forward(self, st: Tensor, labels: Tensor):
        """model.py:81-99 -> (u [B], v [B,S]).  With autograd enabled the parameter gradients are produced by
        csrc/style_train.cu (st and labels are data: no gradient flows into them)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if st.requires_grad or labels.requires_grad:
                raise lib.OsdError('gradients w.r.t. st / labels are not produced (they are data in fit-style)')
            return _StyleFn.apply(self, st.float().contiguous(), labels.float().contiguous(), *self.parameters())
        parr, keep = self._parr()
        return lib.style_forward(parr, st.float().contiguous(), labels.float().contiguous())

    @torch.no_grad()
    def sample(self, labels: Tensor, num_steps: int = 16) -> Tensor:
        """model.py:101-119; the initial noise is drawn exactly like the reference (`th.randn(B, S, device=labels.device)`)."""
        s = torch.randn(labels.size(0), self.style_dim, device=labels.device)
        return self.sample_from(labels, s, num_steps)

    @torch.no_grad()
    def sample_from(self, labels: Tensor, s: Tensor, num_steps: int = 16) -> Tensor:
        parr, keep = self._parr()
        s = s.float().contiguous().clone()
        self.last_eta_u0 = lib.style_sample(parr, labels.float().contiguous(), s, num_steps)
        return s
