"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    python oracle/make_golden.py
Weights and inputs are regenerated from seeds by `oracle/denoiser_oracle.py`
(`make_state_dict(1234)`, `make_inputs(..., seed)`), so only the reference's
OUTPUTS are stored.  Everything here calls the reference's own classes
(`DiffusionModel`, `DiffusionTrainer`, torch AdamW / AveragedModel).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import denoiser_oracle as O  # noqa: E402
from oracle import refimport  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def grad_summary(named_grads):
    names, norms, heads = [], [], []
    for n, g in named_grads:
        names.append(n)
        norms.append(float(g.double().norm()))
        flat = g.reshape(-1)[:8].double().numpy()
        heads.append(np.pad(flat, (0, 8 - flat.size)))
    return np.array(names), np.array(norms), np.stack(heads)


def golden_validation(ns, sd):
    """DiffusionTrainer.validation_step (train.py:128-139) of the unmodified reference on one ragged full-length map:
    l = 8 * 96 + 5 frames -> 8 segments of 96 (the tail is dropped), EMA weights, values read from what it logs."""
    l, seed = 8 * 96 + 5, 41
    inp = O.make_inputs(1, l, seed=seed)
    trainer = ns.DiffusionTrainer(
        val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
        schedule_args=ns.LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
        osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32,
        diffusion_args=refimport.default_args(ns))
    trainer.diffusion_ema.module.load_state_dict(sd)  # the EMA copy is what validation uses; `diffusion` keeps its init
    torch.set_float32_matmul_precision('highest')  # see main(): train.py:53 sets 'medium' process-wide
    labels = torch.zeros(1, 5)
    torch.manual_seed(123)
    trainer.validation_step((inp['h'], inp['x1'], inp['s'], labels), 0)
    logged = {k: float(v) for k, v in trainer.logged.items()}
    torch.manual_seed(123)  # replay the draws of train.py:79,82 at B = 8
    uu = (torch.randperm(8) + torch.rand(8)) / 8
    t = torch.special.ndtri(uu.clamp(1e-6, 1 - 1e-6)).sigmoid()
    # randn_like of the reference's non-contiguous rearranged view (strides (96, 773, 1)) fills a (96, 768, 1)-strided
    # tensor: NOT the values of randn(8, 6, 96) -- replay it on the same view
    from einops import rearrange
    x0 = torch.randn_like(rearrange(inp['x1'][..., :768], '1 ... (b l) -> b ... l', b=8)).contiguous()
    np.savez(os.path.join(OUT, 'val_l773.npz'), l=l, seed=seed, perm_plus_rand=(uu * 8).numpy(), t=t.numpy(), x0=x0.numpy(),
             keys=np.array(sorted(logged)), vals=np.array([logged[k] for k in sorted(logged)]))
    print('validation_step', logged)


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    ns = refimport.import_reference(with_trainer=True)
    sd = O.make_state_dict(1234)
    if len(sys.argv) > 1 and sys.argv[1] == 'val':  # only the round-2 addition (the other fixtures stay byte-identical)
        golden_validation(ns, sd)
        return

    # ---- known-answer constants and default-init facts ----
    m0 = ns.DiffusionModel(6, 128, 32, refimport.default_args(ns)).eval()
    inp = O.make_inputs(2, 64, seed=3)
    with torch.no_grad():
        u_def, v_def = m0(inp['h'], inp['s'], inp['x0'])
    sched = ns.make_lr_schedule(ns.LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000))
    steps = np.array([0, 1, 10, 500, 999, 1000, 1001, 30000, 30001, 60000, 120000])
    np.savez(os.path.join(OUT, 'constants.npz'),
             c0=m0.c0, u_scale=m0.u_scale, u_default=u_def.numpy(), v_default_absmax=float(v_def.abs().max()),
             n_params=sum(p.numel() for p in m0.parameters()),
             lr_steps=steps, lr_vals=np.array([sched(int(s)) for s in steps]),
             names=np.array(list(m0.state_dict().keys())),
             shapes=np.array([str(tuple(v.shape)) for v in m0.state_dict().values()]))

    # ---- forward goldens: fp32 and fp64 reference, B=2, L in {128, 320} ----
    for (B, L, seed) in [(2, 128, 7), (2, 320, 8), (1, 1000, 9)]:
        inp = O.make_inputs(B, L, seed=seed)
        xt = torch.lerp(inp['x0'], inp['x1'], inp['t'][:, None, None])
        m32 = refimport.build_reference_model(ns, sd)
        m64 = refimport.build_reference_model(ns, {k: v.double() for k, v in sd.items()}, torch.float64)
        with torch.no_grad():
            u32, v32 = m32(inp['h'], inp['s'], xt)
            u64, v64 = m64(inp['h'].double(), inp['s'].double(), xt.double())
            # bf16-autocast behaviour of the reference itself (orientation for the bf16 budget)
            with torch.autocast('cpu', dtype=torch.bfloat16):
                ub, vb = m32(inp['h'], inp['s'], xt)
        np.savez(os.path.join(OUT, f'fwd_B{B}_L{L}.npz'), B=B, L=L, seed=seed,
                 u32=u32.numpy(), v32=v32.numpy(), u64=u64.numpy(), v64=v64.numpy(),
                 u_bf16=ub.float().numpy(), v_bf16=vb.float().numpy())
        print('fwd', B, L, 'u', u32.tolist(), 'v absmax', float(v32.abs().max()),
              'bf16 err', float((vb.float() - v64).abs().max() / v64.abs().max()))

    # ---- broadcast-audio (#B = 1) forward, the predict-path shape ----
    inp = O.make_inputs(3, 96, seed=10, a_batch=1)
    m32 = refimport.build_reference_model(ns, sd)
    with torch.no_grad():
        u32, v32 = m32._pred(*m32._precompute_conditioning(inp['h'], inp['s']), inp['x0'])
    np.savez(os.path.join(OUT, 'fwd_bcast_B3_L96.npz'), u32=u32.numpy(), v32=v32.numpy())

    # ---- trainer loss + gradients through the reference's own DiffusionTrainer.forward ----
    B, L, seed = 2, 128, 21
    inp = O.make_inputs(B, L, seed=seed)
    trainer = ns.DiffusionTrainer(
        val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
        schedule_args=ns.LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
        osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32,
        diffusion_args=refimport.default_args(ns))
    trainer.diffusion.load_state_dict(sd)
    # DiffusionTrainer.__init__ sets the process-global float32 matmul precision to 'medium'
    # (train.py:53), which on CPU lets mkldnn run fp32 matmuls in bf16.  The goldens pin the
    # ARITHMETIC, so restore full fp32 before evaluating (documented in DESIGN.md).
    torch.set_float32_matmul_precision('highest')
    labels = torch.zeros(B, 5)
    torch.manual_seed(99)
    loss, logs = trainer(trainer.diffusion, inp['h'], inp['x1'], inp['s'], labels)
    loss.backward()
    # replay the draws (train.py:79,82 order: randperm, rand, randn_like)
    torch.manual_seed(99)
    uu = (torch.randperm(B) + torch.rand(B)) / B
    t = torch.special.ndtri(uu.clamp(1e-6, 1 - 1e-6)).sigmoid()
    x0 = torch.randn_like(inp['x1'])
    names, norms, heads = grad_summary([(n, p.grad) for n, p in trainer.diffusion.named_parameters()])
    np.savez(os.path.join(OUT, f'loss_B{B}_L{L}.npz'), B=B, L=L, seed=seed, t=t.numpy(), x0=x0.numpy(),
             loss=float(loss), osl=float(logs['osl']), del_=float(logs['del']), u_mape=float(logs['u_mape']),
             grad_names=names, grad_norms=norms, grad_heads=heads)
    print('loss', float(loss), {k: float(v) for k, v in logs.items()})

    # ---- sampler: reference DiffusionModel.sample, 8 steps, fp32 ----
    B, L, seed = 2, 128, 31
    inp = O.make_inputs(B, L, seed=seed)
    m32 = refimport.build_reference_model(ns, sd)
    torch.manual_seed(11)
    x_init = torch.randn(B, 6, L)
    torch.manual_seed(11)
    x_fin = m32.sample(inp['h'], inp['s'], 8)
    np.savez(os.path.join(OUT, f'sample_B{B}_L{L}_N8.npz'), B=B, L=L, seed=seed, x_init=x_init.numpy(),
             x_final=x_fin.numpy())
    print('sample absmax', float(x_fin.abs().max()))

    # ---- optimizer + EMA: torch AdamW + AveragedModel on a small tensor, 4 steps ----
    from torch.optim.swa_utils import AveragedModel, get_ema_multi_avg_fn
    g = torch.Generator().manual_seed(5)
    lin = torch.nn.Linear(37, 11)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(11, 37, generator=g))
        lin.bias.copy_(torch.randn(11, generator=g))
    ema = AveragedModel(lin, multi_avg_fn=get_ema_multi_avg_fn(.99))
    opt = torch.optim.AdamW(lin.parameters(), lr=3e-4, weight_decay=0.01)
    lrs = torch.optim.lr_scheduler.LambdaLR(opt, sched)
    p0 = torch.cat([lin.weight.detach().reshape(-1), lin.bias.detach().reshape(-1)]).clone()
    grads, ps, emas, norms_ = [], [], [], []
    for step in range(4):
        gw = torch.randn(11, 37, generator=g) * (3.0 if step == 1 else 0.05)
        gb = torch.randn(11, generator=g) * 0.05
        lin.weight.grad, lin.bias.grad = gw.clone(), gb.clone()
        grads.append(torch.cat([gw.reshape(-1), gb.reshape(-1)]))
        norms_.append(float(torch.nn.utils.clip_grad_norm_(lin.parameters(), 1.0)))
        opt.step()
        lrs.step()
        ema.update_parameters(lin)
        ps.append(torch.cat([lin.weight.detach().reshape(-1), lin.bias.detach().reshape(-1)]).clone())
        emas.append(torch.cat([ema.module.weight.detach().reshape(-1), ema.module.bias.detach().reshape(-1)]).clone())
    np.savez(os.path.join(OUT, 'adamw_ema.npz'), p0=p0.numpy(), grads=torch.stack(grads).numpy(),
             params=torch.stack(ps).numpy(), emas=torch.stack(emas).numpy(), grad_norms=np.array(norms_))
    golden_validation(ns, sd)
    print('done')


if __name__ == '__main__':
    main()
