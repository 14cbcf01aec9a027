"""Helper of tests/test_gpu_nccl.py (launched with torch.distributed.run, one process per GPU, NCCL): one fit-denoiser
training step data-parallel over WORLD ranks vs the same step on the concatenated batch in ONE process.

Every rank holds the full batch of B = 2 * WORLD samples and the draws (t, x0) of the single-process run; rank r trains on
samples [2r, 2r + 2) with its slice of the draws injected, all-reduces the flat gradient over NCCL and applies the fused
clip + AdamW + EMA.  Rank 0 then runs the whole batch alone (world_size = 1) from the same initial weights and compares:
the averaged gradient, its global norm, the updated parameters and the EMA copy.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import denoiser_oracle as O  # noqa: E402  (seeded weights / inputs only)
from osu_dreamer_b200.denoiser import default_args  # noqa: E402
from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs  # noqa: E402


def make_trainer(dev):
    tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                          schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                          osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32, diffusion_args=default_args())
    sd = O.make_state_dict(1234)
    tr.diffusion.load_state_dict(sd)
    tr.diffusion_ema.module.load_state_dict(sd)
    return tr.to(dev)


def step_with_draws(tr, batch, u01, x0, world):
    """training_step with the reference's draws (train.py:79,82) injected: randperm + rand = u01 * B, randn_like = x0"""
    B = x0.shape[0]
    saved = (torch.randperm, torch.rand, torch.randn_like)
    torch.randperm = lambda n, **k: torch.zeros(n, device=k.get('device', 'cpu'))
    torch.rand = lambda n, **k: (u01 * B).to(k.get('device', 'cpu'))
    torch.randn_like = lambda x, **k: x0.to(x.device)
    try:
        return tr.training_step(batch, world_size=world)
    finally:
        torch.randperm, torch.rand, torch.randn_like = saved


def main():
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    L, per = 384, 2
    B = per * world
    inp = {k: v.to(dev) for k, v in O.make_inputs(B, L, seed=17).items()}
    u01 = (torch.randperm(B, generator=torch.Generator().manual_seed(3)) + inp['t'].cpu()) / B  # stratified over the GLOBAL batch
    sl = slice(per * rank, per * rank + per)
    tr = make_trainer(dev)
    lab = torch.zeros(B, 5, device=dev)
    loss, _ = step_with_draws(tr, (inp['h'][sl], inp['x1'][sl], inp['s'][sl], lab[sl]), u01[sl], inp['x0'][sl], world)
    torch.cuda.synchronize()
    g_dp = tr._opt['g'].clone() / world
    losses = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    # replicas stay identical
    p0 = tr._opt['p'].clone()
    dist.broadcast(p0, 0)
    assert torch.equal(p0, tr._opt['p']), f'rank {rank}: replicas diverged after the step'
    if rank == 0:
        one = make_trainer(dev)
        loss1, _ = step_with_draws(one, (inp['h'], inp['x1'], inp['s'], lab), u01, inp['x0'], 1)
        torch.cuda.synchronize()
        g1 = one._opt['g']
        e_loss = abs(float(torch.stack(losses).mean()) - float(loss1)) / abs(float(loss1))
        e_g = float((g_dp - g1).norm() / g1.norm())
        e_norm = abs(float(tr._opt['scal'][0]) - float(one._opt['scal'][0])) / float(one._opt['scal'][0])
        # AdamW's first step is lr * g / (|g| + eps): compare the parameters where the gradient is not at the eps scale
        mask = g1.abs() > 1e-6
        dp = (tr._opt['p'] - one._opt['p']).abs()
        e_p = float(dp[mask].max())
        e_ema = float((tr._opt['ema'] - one._opt['ema']).abs()[mask].max())
        print(f'NCCL world {world}: loss rel err {e_loss:.2e}, gradient rel-L2 {e_g:.2e}, grad-norm rel {e_norm:.2e}, '
              f'max |dp| {e_p:.2e} (lr step {tr.current_lr():.1e}), max |dema| {e_ema:.2e}, compared {int(mask.sum())} of {mask.numel()}', flush=True)
        assert e_loss < 1e-5 and e_g < 1e-4 and e_norm < 1e-5, (e_loss, e_g, e_norm)
        assert e_p < 1e-5 and e_ema < 1e-5, (e_p, e_ema)
        print('NCCL_STEP_OK', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
