"""Autograd plumbing: `DiffusionModel.forward` as one torch.autograd.Function whose forward and backward are
single C-ABI calls (osd_pred_forward(save=1) / osd_pred_backward).  This is what lets the reference's
`DiffusionTrainer.forward` (train.py:69-108) call `model.forward(h, s, xt)` and `loss.backward()` unchanged."""
from __future__ import annotations

import torch

from . import lib


class _DenoiserFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, audio, style, xt, *params):
        if xt.requires_grad or audio.requires_grad or style.requires_grad:
            raise lib.OsdError('gradients w.r.t. audio / style / xt are not produced (they are data in fit-denoiser)')
        audio = audio.float().contiguous()
        style = style.float().contiguous()
        xt = xt.float().contiguous()
        a_tok, cond = model._conditioning_tokens(audio, style)
        u, v = model._pred_tokens(a_tok, cond, audio.shape[0], xt, save=True)
        rt = model._rt
        rt.save_generation = getattr(rt, 'save_generation', 0) + 1
        ctx.model = model
        ctx.generation = rt.save_generation
        ctx.saved = (audio, style, xt, a_tok, cond)
        ctx.shapes = [p.shape for p in params]
        return u, v

    @staticmethod
    def backward(ctx, du, dv):
        model = ctx.model
        rt = model._rt
        if getattr(rt, 'save_generation', 0) != ctx.generation:
            raise lib.OsdError('the saved activations of this forward were overwritten by a later forward of the same '
                               'module (one forward per backward, as in DiffusionTrainer.training_step)')
        audio, style, xt, a_tok, cond = ctx.saved
        B, _, L = xt.shape
        a_batch = audio.shape[0]
        dev = xt.device
        targets = getattr(model, '_grad_targets', None)
        direct = targets is not None and len(targets) == len(ctx.shapes) and targets[0].device == dev
        if direct:
            # the trainer owns one flat gradient buffer: accumulate straight into it, autograd sees no param grads
            grads = targets
        else:
            # one zeroed buffer, every tensor 256-byte aligned (the kernels use 16-byte vector atomics)
            sizes = [int(torch.Size(s).numel()) for s in ctx.shapes]
            padded = [(n + 63) // 64 * 64 for n in sizes]
            flat = torch.zeros(sum(padded), dtype=torch.float32, device=dev)
            grads, off = [], 0
            for s, n, pn in zip(ctx.shapes, sizes, padded):
                grads.append(flat[off:off + n].view(s))
                off += pn
        du = torch.zeros(B, device=dev) if du is None else du.float().contiguous()
        dv = torch.zeros(B, 6, L, device=dev) if dv is None else dv.float().contiguous()
        ws = model._workspace(B, L, a_batch, 1)
        key = ('bwd', B, L, a_batch)
        if key not in rt.ws:
            rt.ws[key] = torch.empty(lib.backward_workspace_bytes(B, L, a_batch), dtype=torch.uint8, device=dev)
        lib.pred_backward(rt.parr, rt.packed, model._mode(), a_tok, cond, model._rope(L, dev), audio, style, xt, du, dv,
                          lib.grad_array(grads), a_batch, ws, rt.ws[key])
        if direct:
            return (None, None, None, None, *([None] * len(grads)))
        return (None, None, None, None, *grads)


def denoiser_apply(model, audio, style, xt):
    model._ensure(xt.device)
    return _DenoiserFn.apply(model, audio, style, xt, *model._params())
