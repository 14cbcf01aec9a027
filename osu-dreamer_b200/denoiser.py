"""`DiffusionModel` -- host-side mirror of osu_dreamer/models/diffusion/model.py:23 of the reference.

Same constructor, attributes (`emb_dim`, `style_dim`, `c0`, `u_scale`), methods
(`_precompute_conditioning`, `_pred`, `forward`, `sample`), parameter names / shapes / initial
distributions and state-dict layout, so it loads reference checkpoints and drops into the reference's
`DiffusionTrainer` / `LDM`.  All tensor math runs in libosd_b200.so (hand-written
sm_100a CUDA) through the C ABI in include/osd_b200.h; torch only owns memory, streams and autograd
plumbing.  There is no CPU fallback: calling the model with non-CUDA tensors raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import Tensor, nn

from . import lib


@dataclass
class BackboneArgs:
    """osu_dreamer/models/diffusion/backbone.py:18-25."""
    depth: int
    expand: int
    head_dim: int
    n_heads: int
    radius: int = 1
    dropout: float = 0.


@dataclass
class DiffusionModelArgs:
    """osu_dreamer/models/diffusion/model.py:16-21."""
    global_cond_dim: int
    backbone_dim: int
    backbone_args: BackboneArgs
    u_head_dim: int = 64


def default_args() -> DiffusionModelArgs:
    """osu_dreamer/models/diffusion/model.yml:77-90."""
    return DiffusionModelArgs(global_cond_dim=512, backbone_dim=512, u_head_dim=64,
                              backbone_args=BackboneArgs(depth=8, expand=4, head_dim=64, n_heads=16, radius=2))


class _Bag(nn.Module):
    """A nameable container: only holds parameters / sub-bags, never called."""


def _parameter_table(E, A, S, Cg, D, H, hd, depth, expand, radius, U):
    """(name, shape, init) in the reference's registration order.  init: 'lin' = torch's default
    Conv1d/Linear init U(+-1/sqrt(fan_in)) for weight and bias, 'zero', 'one', or a float constant."""
    dh = H * hd
    hid = int(D * expand * 2 / 3)
    k = 1 + 2 * radius
    t = [('proj_audio.0', (A, A, 1), 'lin'), ('proj_style.0', (Cg, S), 'lin'), ('proj_in', (D, E, 1), 'lin')]
    for i in range(depth):
        p = f'net.layers.{i}.'
        t += [(p + 'ssg1', (3 * D, Cg), 'zero'), (p + 'proj_cl', (D, A, 1), 'lin'),
              (p + 'attn.qkv_proj', (3 * dh, D, 1), 'lin'), (p + 'attn.out_proj', (D, dh, 1), 'lin'),
              (p + 'attn.q_norm', (hd,), 'one'), (p + 'attn.k_norm', (hd,), 'one'),
              (p + 'ssg2', (3 * D, Cg), 'zero'),
              (p + 'ffn.proj_vg.0', (D, 1, k), 'lin'), (p + 'ffn.proj_vg.1', (2 * hid, D, 1), 'lin'),
              (p + 'ffn.proj_o', (D, hid, 1), 'lin')]
    t += [('proj_out', (E, D, 1), 'zero'),
          ('u_head.0', (E, 1, 3), 'lin'), ('u_head.1', (U, E, 1), 'lin'),
          ('u_head.3', (U, 1, 3), 'lin'), ('u_head.4', (U, U, 1), 'lin'),
          ('u_mod', (2 * U, Cg), 'zero'), ('u_out', (1, U), 'u_out')]
    return t


class _Runtime:
    """Per-module device state (pointer table, packed operand weights, workspaces).  Never copied:
    `copy.deepcopy` (torch's AveragedModel, train.py:67 of the reference) gets a fresh empty one."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.key = None
        self.parr = None
        self.packed = None
        self.ws = {}
        self.rope = {}
        self.graphs = {}
        self.cond_cache = None

    def __deepcopy__(self, memo):
        return _Runtime()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.reset()


class DiffusionModel(nn.Module):
    def __init__(self, emb_dim: int, a_dim: int, style_dim: int, args: DiffusionModelArgs):
        super().__init__()
        ba = args.backbone_args
        if isinstance(ba, dict):
            ba = BackboneArgs(**ba)
        fixed = (emb_dim, a_dim, style_dim, args.global_cond_dim, args.backbone_dim, ba.n_heads, ba.head_dim,
                 ba.depth, ba.expand, ba.radius, args.u_head_dim)
        # the channel widths are compile-time constants of the kernels (model.yml:77-90); the backbone depth is a
        # run-time argument of the C ABI (it rides in `mode`, include/osd_b200.h OSD_MODE)
        if fixed[:7] + fixed[8:] != (6, 128, 32, 512, 512, 16, 64, 4, 2, 64) or not 1 <= ba.depth <= lib.MAX_DEPTH:
            raise ValueError('libosd_b200 is compiled for the reference widths of model.yml:77-90 '
                             f'(6,128,32,512,512,16,64,depth<={lib.MAX_DEPTH},4,2,64); got {fixed}')
        self.depth = ba.depth
        if ba.dropout != 0.:
            raise ValueError('dropout must be 0 (model.yml default); Dropout1d(0.) is the identity')
        self.emb_dim = emb_dim
        self.style_dim = style_dim
        # distance-field constants, model.py:33-43 of the reference
        d0_sq = 2. * emb_dim
        t99 = torch.tensor(2.3263478740408408).sigmoid().item()
        self.c0 = (1 - t99) ** 2 * d0_sq
        self.u_scale = math.sqrt(d0_sq)

        for name, shape, init in _parameter_table(*fixed):
            bag = self
            parts = name.split('.')
            for part in parts:
                if not hasattr(bag, part):
                    bag.add_module(part, _Bag())
                bag = getattr(bag, part)
            if init in ('one',):
                bag.register_parameter('weight', nn.Parameter(torch.ones(shape)))
                continue
            w = torch.empty(shape)
            b = torch.empty(shape[0])
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            if init == 'zero':
                w.zero_(), b.zero_()
            else:
                w.uniform_(-bound, bound), b.uniform_(-bound, bound)
                if init == 'u_out':  # model.py:68-71: zero weight, bias = log(exp(.5) - 1)
                    w.zero_(), b.fill_(-0.4328)
            bag.register_parameter('weight', nn.Parameter(w))
            bag.register_parameter('bias', nn.Parameter(b))
        self.precision = 'bf16'
        self._rt = _Runtime()

    # ------------------------------------------------------------------ runtime plumbing
    def _mode(self):
        """'bf16': bf16 tensor-core operands, fp32 accumulate / residual / statistics (training + inference).
        'fp32': fp32-grade inference -- every tensor-core product runs as the 3-term bf16 split
        a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (tcgen05 has no fp32 MMA and kind::tf32 truncates its operands)."""
        if self.precision == 'bf16':
            return lib.mode_of(lib.MODE_BF16, self.depth)
        if self.precision == 'fp32':
            return lib.mode_of(lib.MODE_F32X3, self.depth)
        raise lib.OsdError(f"unknown precision {self.precision!r}: use 'bf16' or 'fp32'")

    def _params(self):
        return [p for p in self.parameters()]

    def _ensure(self, device):
        """Pointer table + packed operand weights, refreshed when any parameter storage or version changes."""
        rt = self._rt
        ps = self._params()
        if any((not p.is_cuda) for p in ps):
            raise lib.OsdError('DiffusionModel parameters must be on a CUDA device: libosd_b200 has no CPU path')
        key = (self._mode(), tuple(p.data_ptr() for p in ps), tuple(p._version for p in ps))
        if rt.key != key:
            rt.parr = lib.param_array([p.detach() for p in ps])
            need = lib.packed_bytes(self._mode())  # depends on the precision (fp32-grade operands are (hi | lo) pairs)
            if rt.packed is None or rt.packed.device != ps[0].device or rt.packed.numel() != need:
                rt.packed = torch.empty(need, dtype=torch.uint8, device=ps[0].device)
            lib.pack_weights(rt.parr, rt.packed, self._mode())
            rt.key = key
        return rt

    def _workspace(self, B, L, a_batch, save, tag=''):
        rt = self._rt
        k = (B, L, a_batch, save, tag, self._mode())  # sizes depend on the precision mode and the depth
        if k not in rt.ws:
            if len(rt.ws) > 6:
                rt.ws.clear()
            dev = self._params()[0].device
            n = (lib.workspace_bytes(B, L, a_batch, self._mode(), save) if tag == ''
                 else lib.sample_extra_bytes(B, L, a_batch, self._mode()))
            rt.ws[k] = torch.empty(n, dtype=torch.uint8, device=dev)
        return rt.ws[k]

    def _rope(self, L, device):
        rt = self._rt
        if L not in rt.rope:
            if len(rt.rope) > 4:
                rt.rope.clear()
            rt.rope[L] = lib.rope_table(L, device)
        return rt.rope[L]

    def _conditioning_tokens(self, audio: Tensor, style: Tensor):
        """audio [#B,A,l], style [B,S] -> (a_tok [#B*l,128] operand dtype, cond pack fp32)."""
        rt = self._ensure(audio.device)
        a_batch, _, L = audio.shape
        B = style.shape[0]
        mode = self._mode()
        km = 2 if self.precision == 'fp32' else 1  # (hi | lo) bf16 pairs in the fp32-grade mode
        a_tok = torch.empty(a_batch * L, 128 * km, dtype=torch.bfloat16, device=audio.device)
        cond = torch.empty(lib.cond_floats(B, mode), dtype=torch.float32, device=audio.device)
        scratch = torch.empty(a_batch * L * 128, dtype=torch.float32, device=audio.device)
        lib.precompute_conditioning(rt.parr, rt.packed, mode, audio.float().contiguous(), style.float().contiguous(),
                                    scratch, a_tok, cond)
        return a_tok, cond

    # ------------------------------------------------------------------ reference surface
    def _precompute_conditioning(self, audio: Tensor, style: Tensor):
        """model.py:73-84 -> (a [#B,A,l], cg [B,C]) in the reference's channels-first layout."""
        a_tok, cond = self._conditioning_tokens(audio, style)
        a_batch, _, L = audio.shape
        if a_tok.shape[1] == 256:  # fp32-grade mode: a = hi + lo
            a = lib.tokens_to_channels((a_tok[:, :128].float() + a_tok[:, 128:].float()).contiguous(), a_batch, 128, L)
        else:
            a = lib.tokens_to_channels(a_tok, a_batch, 128, L)
        cg = cond[: style.shape[0] * 512].view(style.shape[0], 512)
        # fast path of _pred: if it is handed exactly these tensors, unmodified, the operand copies are reused.  The
        # entry keeps (a, cg) alive, so their addresses cannot be recycled while it exists.
        self._rt.cond_cache = (self._cond_key(a, cg), (a_tok, cond), (a, cg))
        return a, cg

    def _cond_key(self, a: Tensor, cg: Tensor):
        return (self._rt.key, a.data_ptr(), a._version, tuple(a.shape), cg.data_ptr(), cg._version, tuple(cg.shape))

    def _pred(self, a: Tensor, cg: Tensor, xt: Tensor):
        """model.py:86-103 -> (u [B], v [B,E,l]); a function of the (a, cg) it is given, like the reference's: if they
        are not the unmodified tensors `_precompute_conditioning` returned (edited cg, sliced / moved / re-made a), the
        operand copy of `a` and the modulation vectors are rebuilt from their values (osd_conditioning_from)."""
        rt = self._ensure(xt.device)
        cache = getattr(rt, 'cond_cache', None)
        if cache is not None and cache[0] == self._cond_key(a, cg):
            a_tok, cond = cache[1]
        else:
            if a.dim() != 3 or a.shape[1] != 128 or cg.dim() != 2 or cg.shape[1] != 512 or a.shape[0] not in (1, cg.shape[0]):
                raise lib.OsdError(f'_pred: a {tuple(a.shape)} / cg {tuple(cg.shape)} are not [#B,128,l] / [B,512]')
            mode = self._mode()
            km = 2 if self.precision == 'fp32' else 1
            a_tok = torch.empty(a.shape[0] * a.shape[2], 128 * km, dtype=torch.bfloat16, device=xt.device)
            cond = torch.empty(lib.cond_floats(cg.shape[0], mode), dtype=torch.float32, device=xt.device)
            lib.conditioning_from(rt.parr, mode, a.float().contiguous(), cg.float().contiguous(), a.shape[2], a_tok, cond)
        return self._pred_tokens(a_tok, cond, a.shape[0], xt, save=False)

    def _pred_tokens(self, a_tok, cond, a_batch, xt, save):
        rt = self._ensure(xt.device)
        B, _, L = xt.shape
        xt = xt.float().contiguous()
        u = torch.empty(B, dtype=torch.float32, device=xt.device)
        v = torch.empty(B, self.emb_dim, L, dtype=torch.float32, device=xt.device)
        ws = self._workspace(B, L, a_batch, 1 if save else 0)
        lib.pred_forward(rt.parr, rt.packed, self._mode(), a_tok, cond, self._rope(L, xt.device), xt, u, v, a_batch,
                         ws, 1 if save else 0)
        return u, v

    def forward(self, audio: Tensor, style: Tensor, xt: Tensor):
        """model.py:105-114."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if self.precision != 'bf16':
                raise lib.OsdError("training runs in precision='bf16' (the fp32-grade path is inference-only)")
            from .autograd import denoiser_apply
            return denoiser_apply(self, audio, style, xt)
        a_tok, cond = self._conditioning_tokens(audio, style)
        return self._pred_tokens(a_tok, cond, audio.shape[0], xt, save=False)

    @torch.no_grad()
    def sample(self, audio: Tensor, style: Tensor, num_steps: int, show_progress: bool = False) -> Tensor:
        """model.py:117-138.  The initial noise is drawn exactly like the reference (global generator,
        `th.randn(B, E, l, device=audio.device)`); the whole (num_steps + 1)-forward loop is one C-ABI
        call with the step-invariant proj_cl hoisted and no host synchronisation."""
        x = torch.randn(style.size(0), self.emb_dim, audio.size(-1), device=audio.device)
        return self.sample_from(audio, style, x, num_steps)

    #: sampler launch mode: True = always replay a captured CUDA graph, False = never, None = automatic (graphs for
    #: launch-bound shapes: B * l <= GRAPH_AUTO_TOKENS, e.g. `predict` on one song -- l ~ 1-2 k frames, B = a few
    #: difficulties -- where the ~6500 kernels of a 64-step sample are each only microseconds long)
    graph_sampler = None
    GRAPH_AUTO_TOKENS = 32768

    def _sample_eager(self, audio: Tensor, style: Tensor, x: Tensor, num_steps: int, eta_u0: Tensor):
        """conditioning + the whole (num_steps + 1)-forward loop, in place on x; enqueues only (capturable)."""
        rt = self._rt
        a_tok, cond = self._conditioning_tokens(audio, style)
        a_batch, _, L = audio.shape
        B = style.shape[0]
        ws = self._workspace(B, L, a_batch, 0)
        extra = self._workspace(B, L, a_batch, 0, tag='sample')
        rope = self._rope(L, x.device)
        lib.sample(rt.parr, rt.packed, self._mode(), a_tok, cond, rope, x, num_steps, float(self.c0), a_batch, ws,
                   extra, eta_u0)
        return (a_tok, cond, ws, extra, rope, rt.parr, rt.packed)  # everything the enqueued work points at

    @torch.no_grad()
    def sample_from(self, audio: Tensor, style: Tensor, x: Tensor, num_steps: int) -> Tensor:
        rt = self._ensure(audio.device)
        a_batch, _, L = audio.shape
        B = style.shape[0]
        use_graph = self.graph_sampler if self.graph_sampler is not None else (B * L <= self.GRAPH_AUTO_TOKENS)
        if not use_graph:
            x = x.float().contiguous().clone()
            self.last_eta_u0 = torch.empty(2, dtype=torch.float32, device=x.device)
            self._sample_eager(audio.float().contiguous(), style.float().contiguous(), x, num_steps, self.last_eta_u0)
            return x
        # CUDA-graph path: first call of a shape runs eagerly (lazy module loading, one-time function attributes,
        # workspace / rope allocation), the second captures, later ones only copy the inputs and replay.
        key = (rt.key, B, L, a_batch, int(num_steps), audio.device)
        ent = rt.graphs.get(key)
        if ent is None:
            if len(rt.graphs) > 4:
                rt.graphs.clear()
            x = x.float().contiguous().clone()
            self.last_eta_u0 = torch.empty(2, dtype=torch.float32, device=x.device)
            self._sample_eager(audio.float().contiguous(), style.float().contiguous(), x, num_steps, self.last_eta_u0)
            rt.graphs[key] = 'warm'
            return x
        if ent == 'warm':
            dev = audio.device
            st = {'audio': torch.empty(a_batch, audio.shape[1], L, dtype=torch.float32, device=dev),
                  'style': torch.empty(B, style.shape[1], dtype=torch.float32, device=dev),
                  'x': torch.empty(B, self.emb_dim, L, dtype=torch.float32, device=dev),
                  'eta_u0': torch.empty(2, dtype=torch.float32, device=dev)}
            st['audio'].copy_(audio)
            st['style'].copy_(style)
            st['x'].copy_(x)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st['keep'] = self._sample_eager(st['audio'], st['style'], st['x'], num_steps, st['eta_u0'])
            st['graph'] = g
            ent = rt.graphs[key] = st
        ent['audio'].copy_(audio)
        ent['style'].copy_(style)
        ent['x'].copy_(x)
        ent['graph'].replay()
        self.last_eta_u0 = ent['eta_u0'].clone()
        return ent['x'].clone()


Denoiser = DiffusionModel  # the north-star's name for the same module
