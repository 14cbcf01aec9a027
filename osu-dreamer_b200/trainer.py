"""`DiffusionTrainer` -- host-side mirror of osu_dreamer/models/diffusion/train.py:33 of the reference,
without pytorch_lightning (not installed here): same constructor arguments, same loss, same optimizer /
LR schedule / gradient clipping / EMA semantics, same state-dict keys (`diffusion.*`,
`diffusion_ema.module.*`, `diffusion_ema.n_averaged`) so checkpoints interchange with the reference's
`export-inference` (osu_dreamer/models/inference/artifact.py:18-42).

Data-parallel training (`fit-denoiser` on N GPUs of one box): one process per GPU, a full replica per
rank, one NCCL all-reduce (sum) over the flat gradient buffer after backward, then the fused
clip + AdamW + EMA kernel (rank-local, identical on every rank).  The EMA copy is never part of the
all-reduce (hard part 6).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any

import torch
from torch import Tensor, nn

from . import lib
from .denoiser import DiffusionModel, DiffusionModelArgs, BackboneArgs


@dataclass(kw_only=True)
class LRScheduleArgs:
    """osu_dreamer/common/lr_schedule.py:4-8."""
    warmup_steps: int = 0
    warmup_init: float = 1
    decay_start: float = float('inf')


def make_lr_schedule(lr: LRScheduleArgs):
    """osu_dreamer/common/lr_schedule.py:10-22: exponential warm-up, inverse-sqrt decay."""
    assert lr.warmup_steps <= lr.decay_start

    def schedule(step: int) -> float:
        if step < lr.warmup_steps:
            return lr.warmup_init ** (1 - step / lr.warmup_steps)
        if step > lr.decay_start:
            return (step / lr.decay_start) ** -.5
        return 1.

    return schedule


class _FusedLoss(torch.autograd.Function):
    """loss(u_pred, v_pred) with the value and both gradients produced by one fused CUDA pass (osd_loss_fwd_bwd)."""

    @staticmethod
    def forward(ctx, u_pred, v_pred, xt, x1, c0, osl_w, del_w):
        out4, du, dv = lib.loss_fwd_bwd(xt.float().contiguous(), x1.float().contiguous(), u_pred.float().contiguous(),
                                        v_pred.float().contiguous(), c0, osl_w, del_w)
        ctx.save_for_backward(du, dv)
        ctx.mark_non_differentiable(out4)
        return out4[0].clone(), out4

    @staticmethod
    def backward(ctx, g, _g4):
        du, dv = ctx.saved_tensors
        return du * g, dv * g, None, None, None, None, None


def frame_dist_sq(a: Tensor, b: Tensor) -> Tensor:
    """train.py:22-31: squared distance in the per-frame metric (sum over channels, mean over length)."""
    return (a - b).square().sum(1).mean(1)


class _EMA(nn.Module):
    """State-dict compatible stand-in for torch.optim.swa_utils.AveragedModel (`module`, `n_averaged`)."""

    def __init__(self, model: DiffusionModel):
        super().__init__()
        import copy
        self.module = copy.deepcopy(model)
        self.register_buffer('n_averaged', torch.tensor(0, dtype=torch.long))


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _flatten(params):
    """Re-home the parameters as views of one flat fp32 buffer (order preserved, every tensor 256-byte
    aligned because the kernels use 16-byte vector loads / atomics; the gaps stay zero) and return it."""
    params = list(params)
    n = sum(_pad64(p.numel()) for p in params)
    flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        k = p.numel()
        flat[off:off + k].copy_(p.data.reshape(-1))
        p.data = flat[off:off + k].view(p.shape)
        off += _pad64(k)
    return flat


def _is_flat(params, flat):
    if flat is None:
        return False
    off = flat.data_ptr()
    for p in params:
        if p.data_ptr() != off or p.device != flat.device:
            return False
        off += _pad64(p.numel()) * 4
    return True


class DiffusionTrainer(nn.Module):
    def __init__(self, val_batches: int, opt_args: dict[str, Any], schedule_args: LRScheduleArgs, osl_weight: float,
                 del_weight: float, emb_dim: int, a_dim: int, style_dim: int, diffusion_args: DiffusionModelArgs,
                 gradient_clip_val: float = 1.0):
        super().__init__()
        if isinstance(schedule_args, dict):
            schedule_args = LRScheduleArgs(**schedule_args)
        if isinstance(diffusion_args, dict):
            d = dict(diffusion_args)
            if isinstance(d.get('backbone_args'), dict):
                d['backbone_args'] = BackboneArgs(**d['backbone_args'])
            diffusion_args = DiffusionModelArgs(**d)
        self.hparams = dict(val_batches=val_batches, opt_args=opt_args, schedule_args=schedule_args,
                            osl_weight=osl_weight, del_weight=del_weight, emb_dim=emb_dim, a_dim=a_dim,
                            style_dim=style_dim, diffusion_args=diffusion_args)
        self.val_batches = val_batches
        self.opt_args = dict(opt_args)
        self.lr_schedule = make_lr_schedule(schedule_args)
        self.osl_weight = osl_weight
        self.del_weight = del_weight
        self.gradient_clip_val = gradient_clip_val  # trainer.gradient_clip_val in model.yml:39
        self.diffusion = DiffusionModel(emb_dim, a_dim, style_dim, diffusion_args)
        self.diffusion_ema = _EMA(self.diffusion)
        self.global_step = 0  # optimizer steps taken: drives the LR schedule (Lightning's global_step)
        self.adam_step = 0    # steps the AdamW moments have seen: drives the bias corrections; differs from global_step after
        #                       resuming from a checkpoint without optimizer state (cli.load_checkpoint)
        self._opt = None  # lazily built flat optimizer state
        self._ema_updates = None

    # ------------------------------------------------------------------ loss (train.py:69-108)
    def forward(self, model: DiffusionModel, h: Tensor, x1: Tensor, s: Tensor, _labels: Tensor | None = None):
        B = x1.size(0)
        # stratified logit-normal noise; draw order randperm, rand, randn_like as in train.py:79,82
        u = (torch.randperm(B, device=x1.device) + torch.rand(B, device=x1.device)) / B
        t = torch.special.ndtri(u.clamp(1e-6, 1 - 1e-6)).sigmoid().to(x1.dtype)
        x0 = torch.randn_like(x1)
        xt = torch.lerp(x0, x1, t[:, None, None])
        u_pred, v_pred = model.forward(h, s, xt)

        if x1.is_cuda:  # fused loss + output gradients (the torch expression below is its specification)
            loss, out4 = _FusedLoss.apply(u_pred, v_pred, xt, x1, float(model.c0), float(self.osl_weight),
                                          float(self.del_weight))
            return loss, {'loss': out4[0], 'osl': out4[1], 'del': out4[2], 'u_mape': out4[3]}

        d_sq = frame_dist_sq(xt, x1)
        u_target = (d_sq + model.c0).sqrt()
        denoised = xt - u_pred[:, None, None] * v_pred
        osl = (frame_dist_sq(denoised, x1) / (d_sq + model.c0)).mean()
        v_target = (xt - x1) / u_target[:, None, None]
        del_ = frame_dist_sq(v_pred, v_target).mean()
        loss = self.osl_weight * osl + self.del_weight * del_
        u_err = ((u_pred - u_target) / u_target).abs().mean()
        return loss, {'loss': loss.detach(), 'osl': osl.detach(), 'del': del_.detach(), 'u_mape': u_err.detach()}

    # ------------------------------------------------------------------ optimizer state
    def configure_optimizers(self):
        """Flat-buffer AdamW state (train.py:110-118).  Only `self.diffusion.*` is optimised; the reference's
        AdamW also receives the EMA copy's parameters but they never get gradients, so they are never stepped."""
        model_params = list(self.diffusion.parameters())
        ema_params = list(self.diffusion_ema.module.parameters())
        o = self._opt or {}
        if not _is_flat(model_params, o.get('p')):
            o['p'] = _flatten(model_params)
            dev, n = o['p'].device, o['p'].numel()
            o['g'] = torch.zeros(n, dtype=torch.float32, device=dev)
            o['m'] = torch.zeros(n, dtype=torch.float32, device=dev) if 'm' not in o or o['m'].numel() != n else o['m'].to(dev)
            o['v'] = torch.zeros(n, dtype=torch.float32, device=dev) if 'v' not in o or o['v'].numel() != n else o['v'].to(dev)
            o['acc'] = torch.zeros(1, dtype=torch.float64, device=dev)
            o['scal'] = torch.zeros(2, dtype=torch.float32, device=dev)
            off, targets = 0, []
            for p in model_params:
                k = p.numel()
                targets.append(o['g'][off:off + k].view(p.shape))
                off += _pad64(k)
            o['targets'] = targets  # handed to the model only for the duration of training_step (direct-gradient mode)
        if not _is_flat(ema_params, o.get('ema')):
            o['ema'] = _flatten(ema_params)
        self._opt = o
        return o

    def zero_grad(self, set_to_none: bool = True):
        if self._opt is not None:
            self._opt['g'].zero_()
        for p in self.diffusion.parameters():
            p.grad = None

    def current_lr(self) -> float:
        return self.opt_args.get('lr', 1e-3) * self.lr_schedule(self.global_step)

    def optimizer_step(self, world_size: int = 1):
        """grad all-reduce (when distributed) + fused clip/AdamW/EMA; mirrors Lightning's order:
        backward -> clip -> optimizer.step -> lr_scheduler.step -> on_train_batch_end (EMA update)."""
        o = self.configure_optimizers()
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(o['g'])  # NCCL sum over NVLink; scaled by 1/world inside the fused kernel
        betas = self.opt_args.get('betas', (0.9, 0.999))
        if self._ema_updates is None:  # one-time host read (e.g. after loading a checkpoint)
            self._ema_updates = int(self.diffusion_ema.n_averaged.item())
        lib.adamw_ema_step(o['p'], o['g'], o['m'], o['v'], o['ema'], self.adam_step + 1, self.current_lr(), betas[0],
                           betas[1], self.opt_args.get('eps', 1e-8), self.opt_args.get('weight_decay', 1e-2),
                           self.gradient_clip_val or 0.0, 1.0 / world_size, 0.99, self._ema_updates == 0,
                           o['acc'], o['scal'])
        self._ema_updates += 1
        self.diffusion_ema.n_averaged += 1
        self.global_step += 1
        self.adam_step += 1
        # parameters changed in place through the flat buffer: invalidate the packed operand copies
        self.diffusion._rt.key = None
        self.diffusion_ema.module._rt.key = None

    def training_step(self, batch, batch_idx: int = 0, world_size: int = 1):
        """train.py:120-126 + the Lightning loop body around it (model.yml: clip 1.0, accumulate 1)."""
        o = self.configure_optimizers()
        self.zero_grad()
        # direct-gradient mode, scoped to this call: the backward of `diffusion` accumulates straight into the flat buffer and
        # autograd sees no parameter gradients; any other autograd user of the module (torch.autograd.grad, .grad readers)
        # gets ordinary gradients
        self.diffusion._grad_targets = o['targets']
        try:
            loss, log = self(self.diffusion, *batch)
            loss.backward()
        finally:
            self.diffusion._grad_targets = None
        self.optimizer_step(world_size)
        return loss.detach(), log

    @torch.no_grad()
    def validation_step(self, batch, batch_idx: int = 0):
        """train.py:128-139: one full-length map cut into `val_batches` segments, EMA weights."""
        h, z, s, l = batch
        seg = z.size(-1) // self.val_batches
        bl = self.val_batches * seg
        nb = self.val_batches
        h = h[..., :bl].reshape(h.shape[1], nb, seg).permute(1, 0, 2).contiguous()
        z = z[..., :bl].reshape(z.shape[1], nb, seg).permute(1, 0, 2).contiguous()
        s = s.expand(nb, -1).contiguous()
        l = l.expand(nb, -1).contiguous()
        _, log = self(self.diffusion_ema.module, h, z, s, l)
        return {f'val/{k}': v for k, v in log.items()}
