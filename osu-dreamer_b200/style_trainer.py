"""`StyleTrainer` -- host-side mirror of osu_dreamer/models/style/train.py:16-160 of the reference without
pytorch_lightning: same constructor arguments, same loss and draw order (randperm, rand, randn_like, rand_like --
train.py:59-66), same optimizer (AdamW over `style.*` only, train.py:93-94) / LR schedule / clip (model.yml:33) / EMA
(`style_ema`, get_ema_multi_avg_fn(.99), train.py:46,109) and the same state-dict keys (`style.*`, `style_ema.module.*`,
`style_ema.n_averaged`), so its checkpoints feed the reference's `export-inference` (models/inference/artifact.py:31-35).

The model's forward / backward run in csrc/style_train.cu, the loss in `osd_style_loss`, the optimizer tail in the same
fused clip + AdamW + EMA kernel as fit-denoiser.  Data-parallel like fit-denoiser (one NCCL all-reduce of the flat
gradient) although at 6 M parameters and batch 512 one GPU is the sensible configuration (model.yml:9).
"""
from __future__ import annotations

import copy
from typing import Any

import torch
from torch import Tensor, nn

from . import lib
from .style import StyleModel, StyleModelArgs
from .trainer import LRScheduleArgs, _flatten, _is_flat, _pad64, make_lr_schedule


class _StyleLoss(torch.autograd.Function):
    """loss(u_pred, v_pred) of train.py:70-88 with value and both output gradients from one fused pass (osd_style_loss)."""

    @staticmethod
    def forward(ctx, u_pred, v_pred, st, s1, osl_w, del_w):
        out4, du, dv = lib.style_loss(st.float().contiguous(), s1.float().contiguous(), u_pred.float().contiguous(),
                                      v_pred.float().contiguous(), osl_w, del_w)
        ctx.save_for_backward(du, dv)
        ctx.mark_non_differentiable(out4)
        return out4[0].clone(), out4

    @staticmethod
    def backward(ctx, g, _g4):
        du, dv = ctx.saved_tensors
        return du * g, dv * g, None, None, None, None


class _EMA(nn.Module):
    """State-dict compatible stand-in for torch.optim.swa_utils.AveragedModel (`module`, `n_averaged`)."""

    def __init__(self, model: StyleModel):
        super().__init__()
        self.module = copy.deepcopy(model)
        self.register_buffer('n_averaged', torch.tensor(0, dtype=torch.long))


def energy_distance(x: Tensor, y: Tensor) -> Tensor:
    """train.py:153-160."""
    def mean_dist(a, b, exclude_diag: bool):
        d = torch.cdist(a, b)
        if exclude_diag:
            n = a.size(0)
            return d.sum() / (n * (n - 1))
        return d.mean()
    return 2 * mean_dist(x, y, False) - mean_dist(x, x, True) - mean_dist(y, y, True)


class StyleTrainer(nn.Module):
    def __init__(self, opt_args: dict[str, Any], schedule_args: LRScheduleArgs, label_drop_prob: float, osl_weight: float,
                 del_weight: float, style_dim: int, style_args: StyleModelArgs, gradient_clip_val: float = 1.0):
        super().__init__()
        if isinstance(schedule_args, dict):
            schedule_args = LRScheduleArgs(**schedule_args)
        if isinstance(style_args, dict):
            style_args = StyleModelArgs(**style_args)
        self.hparams = dict(opt_args=opt_args, schedule_args=schedule_args, label_drop_prob=label_drop_prob,
                            osl_weight=osl_weight, del_weight=del_weight, style_dim=style_dim, style_args=style_args)
        self.opt_args = dict(opt_args)
        self.lr_schedule = make_lr_schedule(schedule_args)
        self.label_drop_prob = label_drop_prob
        self.osl_weight = osl_weight
        self.del_weight = del_weight
        self.gradient_clip_val = gradient_clip_val  # trainer.gradient_clip_val in models/style/model.yml:33
        self.style = StyleModel(style_dim, style_args)
        self.style_ema = _EMA(self.style)
        self.global_step = 0
        self.adam_step = 0
        self._opt = None
        self._ema_updates = None
        self._val_s: list[Tensor] = []
        self._val_labels: list[Tensor] = []

    # ------------------------------------------------------------------ loss (train.py:48-91)
    def forward(self, model: StyleModel, _h: Tensor, _z: Tensor, s1: Tensor, labels: Tensor):
        B = s1.size(0)
        u = (torch.randperm(B, device=s1.device) + torch.rand(B, device=s1.device)) / B
        t = torch.special.ndtri(u.clamp(1e-6, 1 - 1e-6)).sigmoid().to(s1.dtype)
        s0 = torch.randn_like(s1)
        st = torch.lerp(s0, s1, t[:, None])
        masked_labels = torch.where(torch.rand_like(labels) < self.label_drop_prob, -1, labels)
        u_pred, v_pred = model(st, masked_labels)
        loss, out4 = _StyleLoss.apply(u_pred, v_pred, st, s1, float(self.osl_weight), float(self.del_weight))
        return loss, {'loss': out4[0], 'osl': out4[1], 'del': out4[2], 'u_mape': out4[3]}

    # ------------------------------------------------------------------ optimizer state (train.py:93-109)
    def configure_optimizers(self):
        model_params = list(self.style.parameters())
        ema_params = list(self.style_ema.module.parameters())
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
                targets.append(o['g'][off:off + p.numel()].view(p.shape))
                off += _pad64(p.numel())
            o['targets'] = targets
        if not _is_flat(ema_params, o.get('ema')):
            o['ema'] = _flatten(ema_params)
        self._opt = o
        return o

    def zero_grad(self, set_to_none: bool = True):
        if self._opt is not None:
            self._opt['g'].zero_()
        for p in self.style.parameters():
            p.grad = None

    def current_lr(self) -> float:
        return self.opt_args.get('lr', 1e-3) * self.lr_schedule(self.global_step)

    def optimizer_step(self, world_size: int = 1):
        o = self.configure_optimizers()
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(o['g'])
        betas = self.opt_args.get('betas', (0.9, 0.999))
        if self._ema_updates is None:
            self._ema_updates = int(self.style_ema.n_averaged.item())
        lib.adamw_ema_step(o['p'], o['g'], o['m'], o['v'], o['ema'], self.adam_step + 1, self.current_lr(), betas[0], betas[1],
                           self.opt_args.get('eps', 1e-8), self.opt_args.get('weight_decay', 1e-2), self.gradient_clip_val or 0.0,
                           1.0 / world_size, 0.99, self._ema_updates == 0, o['acc'], o['scal'])
        self._ema_updates += 1
        self.style_ema.n_averaged += 1
        self.global_step += 1
        self.adam_step += 1

    def training_step(self, batch, batch_idx: int = 0, world_size: int = 1):
        """train.py:102-109 + the Lightning loop body around it (clip 1.0, accumulate 1)."""
        o = self.configure_optimizers()
        self.zero_grad()
        self.style._grad_targets = o['targets']  # direct-gradient mode for the duration of this call
        try:
            loss, log = self(self.style, *batch)
            loss.backward()
        finally:
            self.style._grad_targets = None
        self.optimizer_step(world_size)
        return loss.detach(), log

    # ------------------------------------------------------------------ validation (train.py:111-150)
    def on_validation_epoch_start(self):
        self._val_s, self._val_labels = [], []

    def validation_step(self, batch, batch_idx: int = 0):
        _, _, s, labels = batch
        self._val_s.append(s.detach())
        self._val_labels.append(labels.detach())

    @torch.no_grad()
    def on_validation_epoch_end(self) -> dict:
        s_real, labels = torch.cat(self._val_s), torch.cat(self._val_labels)
        B = s_real.size(0)
        ema = self.style_ema.module
        e = torch.empty(B, 0, 0, device=s_real.device)
        _, log = self(ema, e, e, s_real, labels)
        out = {f'val/{k}': v for k, v in log.items()}
        if B < 2:
            return out
        K = 4
        samp = torch.stack([ema.sample(labels, 16) for _ in range(K)])  # K B S
        rr = torch.cdist(s_real, s_real).fill_diagonal_(torch.inf).min(1).values.mean()
        flat = samp.flatten(0, 1)
        out['val/nn_ratio'] = torch.cdist(flat, s_real).min(1).values.mean() / rr
        hi = labels[:, 0] >= 5
        if hi.sum() > 1:
            R = s_real[hi]
            rr_hi = torch.cdist(R, R).fill_diagonal_(torch.inf).min(1).values.mean()
            out['val/nn_ratio_sr5'] = torch.cdist(samp[:, hi].flatten(0, 1), R).min(1).values.mean() / rr_hi
        out['val/cond_recall'] = (samp - s_real[None]).norm(dim=-1).min(0).values.mean()
        per_cond = samp.transpose(0, 1)
        out['val/sample_spread'] = torch.cdist(per_cond, per_cond).sum() / (K * (K - 1) * per_cond.size(0)) / rr
        out['val/energy_dist'] = energy_distance(flat, s_real)
        return out
