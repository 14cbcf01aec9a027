"""GPU: fused clip + AdamW + EMA kernel vs the golden produced by torch.optim.AdamW / AveragedModel /
clip_grad_norm_ (the reference's optimizer stack), and an end-to-end trainer smoke (loss decreases)."""
import os

import numpy as np
import pytest
import torch

from oracle import denoiser_oracle as O

pytestmark = pytest.mark.gpu


def test_fused_adamw_ema_matches_torch_golden(golden_dir):
    from osu_dreamer_b200 import lib
    g = np.load(os.path.join(golden_dir, 'adamw_ema.npz'))
    p = torch.from_numpy(g['p0']).cuda()
    n = p.numel()
    m, v, ema = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    scal = torch.zeros(2, device='cuda')
    for step in range(4):
        grad = torch.from_numpy(g['grads'][step]).cuda()
        lr = 3e-4 * O.lr_lambda(step)
        lib.adamw_ema_step(p, grad, m, v, ema, step + 1, lr, 0.9, 0.999, 1e-8, 0.01, 1.0, 1.0, 0.99, step == 0, acc, scal)
        torch.cuda.synchronize()
        assert abs(float(scal[0]) - g['grad_norms'][step]) < 1e-4 * g['grad_norms'][step]
        assert np.allclose(p.cpu().numpy(), g['params'][step], rtol=2e-5, atol=2e-6)
        assert np.allclose(ema.cpu().numpy(), g['emas'][step], rtol=2e-5, atol=2e-6)


def test_fused_loss_matches_torch_expression():
    """osd_loss_fwd_bwd vs the reference's torch expression (train.py:86-101) and its autograd gradients."""
    from osu_dreamer_b200 import lib
    B, L = 3, 700
    g = torch.Generator().manual_seed(3)
    xt, x1 = torch.randn(B, 6, L, generator=g).cuda(), torch.randn(B, 6, L, generator=g).cuda()
    u = (torch.rand(B, generator=g) * 3 + 0.2).cuda().requires_grad_(True)
    v = torch.randn(B, 6, L, generator=g).cuda().requires_grad_(True)
    c0, ow, dw = 0.0949756, 1.0, 30.0
    d_sq = O.frame_dist_sq(xt, x1)
    ut = (d_sq + c0).sqrt()
    osl = (O.frame_dist_sq(xt - u[:, None, None] * v, x1) / (d_sq + c0)).mean()
    del_ = O.frame_dist_sq(v, (xt - x1) / ut[:, None, None]).mean()
    loss = ow * osl + dw * del_
    loss.backward()
    mape = ((u - ut) / ut).abs().mean()
    out4, du, dv = lib.loss_fwd_bwd(xt, x1, u.detach(), v.detach(), c0, ow, dw)
    torch.cuda.synchronize()
    ref = torch.stack([loss, osl, del_, mape]).detach()
    assert torch.allclose(out4, ref, rtol=2e-5, atol=1e-6), (out4, ref)
    assert torch.allclose(du, u.grad, rtol=1e-4, atol=1e-7)
    assert float((dv - v.grad).abs().max()) <= 1e-5 * float(v.grad.abs().max())


def test_trainer_steps_and_state_dict_keys(oracle_sd):
    from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs
    from osu_dreamer_b200.denoiser import default_args
    torch.manual_seed(0)
    tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                          schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                          osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32,
                          diffusion_args=default_args())
    tr.diffusion.load_state_dict(oracle_sd)
    tr.diffusion_ema.module.load_state_dict(oracle_sd)
    tr = tr.cuda()
    keys = list(tr.state_dict().keys())
    assert keys[0] == 'diffusion.proj_audio.0.weight' and 'diffusion_ema.module.u_out.bias' in keys
    assert 'diffusion_ema.n_averaged' in keys and len(keys) == 2 * 164 + 1
    inp = O.make_inputs(4, 256, seed=5)
    batch = (inp['h'].cuda(), inp['x1'].cuda(), inp['s'].cuda(), torch.zeros(4, 5).cuda())
    p_before = tr.diffusion.proj_in.weight.detach().clone()
    losses = []
    for i in range(6):
        torch.manual_seed(100)  # same noise draw every step -> the loss must go down
        loss, log = tr.training_step(batch)
        losses.append(float(loss))
    torch.cuda.synchronize()
    print('losses', losses, 'grad norm', float(tr._opt['scal'][0]))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert not torch.equal(p_before, tr.diffusion.proj_in.weight.detach())
    assert int(tr.diffusion_ema.n_averaged) == 6 and tr.global_step == 6
    # EMA tracks the parameters: after the first update it equals them, later it lags
    d = (tr.diffusion_ema.module.proj_in.weight - tr.diffusion.proj_in.weight).abs().max()
    assert float(d) > 0
    val = tr.validation_step((inp['h'][:1].cuda(), inp['x1'][:1].cuda(), inp['s'][:1].cuda(), torch.zeros(1, 5).cuda()))
    assert set(val) == {'val/loss', 'val/osl', 'val/del', 'val/u_mape'}


def test_first_step_matches_oracle_adamw(oracle_sd):
    """one full training step (CUDA) vs the oracle: same noise, oracle grads (CPU autograd) + oracle AdamW."""
    from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs
    from osu_dreamer_b200.denoiser import default_args
    tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                          schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                          osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32,
                          diffusion_args=default_args())
    tr.diffusion.load_state_dict(oracle_sd)
    tr = tr.cuda()
    B, L = 2, 128
    inp = O.make_inputs(B, L, seed=21)
    batch = (inp['h'].cuda(), inp['x1'].cuda(), inp['s'].cuda(), torch.zeros(B, 5).cuda())
    torch.manual_seed(99)
    tr.training_step(batch)
    torch.cuda.synchronize()
    # replay the CUDA generator draws for the oracle
    torch.manual_seed(99)
    u = (torch.randperm(B, device='cuda') + torch.rand(B, device='cuda')) / B
    t = torch.special.ndtri(u.clamp(1e-6, 1 - 1e-6)).sigmoid().cpu()
    x0 = torch.randn(B, 6, L, device='cuda').cpu()
    sd = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    loss, _ = O.trainer_loss(sd, inp['h'], inp['x1'], inp['s'], x0, t)
    loss.backward()
    gn = torch.sqrt(sum(v.grad.double().pow(2).sum() for v in sd.values())).item()
    coef = min(1.0, 1.0 / (gn + 1e-6))
    assert abs(float(tr._opt['scal'][0]) - gn) < 3e-2 * gn
    worst = 0.0
    for name, p in tr.diffusion.named_parameters():
        ref = sd[name].detach().clone()
        m, v, ema = torch.zeros_like(ref), torch.zeros_like(ref), torch.zeros_like(ref)
        O.adamw_ema_step(ref, sd[name].grad, m, v, ema, 1, 3e-4 * O.lr_lambda(0), clip_coef=coef, ema_first=True)
        # first Adam step moves every element by ~lr * sign(g): compare the update direction where |g| is not tiny
        upd, upd_ref = (p.detach().cpu() - oracle_sd[name]), (ref - oracle_sd[name])
        mask = sd[name].grad.abs() > 1e-3 * sd[name].grad.abs().max()
        agree = float((torch.sign(upd[mask]) == torch.sign(upd_ref[mask])).float().mean()) if mask.any() else 1.0
        worst = max(worst, 1 - agree)
    print('worst sign disagreement of the first update', worst)
    assert worst < 0.05


@pytest.mark.gpu
def test_device_feeder_overlapped_h2d_matches_host_batches(tmp_path):
    """DeviceFeeder: pinned ring + side-stream H2D; the device batches must equal the host collation, in order, also
    when the consumer is slower or faster than the producer and the ring wraps several times"""
    import time
    import numpy as np
    from osu_dreamer_b200.data import DeviceFeeder, LatentWindows, batches
    rng = np.random.default_rng(0)
    for ms in range(6):
        d = tmp_path / f'set{ms}'
        d.mkdir()
        l = 900 + 41 * ms
        np.save(d / 'h.npy', rng.standard_normal((128, l)).astype(np.float32))
        for k in range(3):
            np.savez(d / f'm{k}.latent.npz', z=rng.standard_normal((6, l)).astype(np.float32),
                     s=rng.standard_normal(32).astype(np.float32), labels=rng.random(5).astype(np.float32))
    sets = sorted(p for p in tmp_path.iterdir() if p.is_dir())
    direct = list(batches(LatentWindows(sets, 96, 8, -1, seed=5), 4, pin=False))
    assert len(direct) >= 12
    for delay in (0.0, 0.01):
        got = []
        for b in DeviceFeeder(LatentWindows(sets, 96, 8, -1, seed=5), 4, device='cuda', depth=2):
            assert all(t.is_cuda for t in b)
            got.append(tuple((t * 1.0).cpu() for t in b))  # consume on the current stream
            time.sleep(delay)
        assert len(got) == len(direct)
        for a, b in zip(direct, got):
            assert all(torch.equal(x, y) for x, y in zip(a, b))
