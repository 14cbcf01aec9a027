"""CPU, world_size 2, gloo: the data-parallel host logic of fit-denoiser -- per-rank data shards, summing
all-reduce of the flat gradient buffer, 1/world scaling + clip + AdamW identical on every rank.
(The fused CUDA optimizer kernel is checked against the same oracle on the GPU in test_gpu_trainer.py.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import denoiser_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, cache):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from osu_dreamer_b200.data import synthetic_batches
    from osu_dreamer_b200.trainer import _flatten, _pad64
    torch.manual_seed(0)
    # identical replicas
    ps = [torch.nn.Parameter(torch.randn(37, 11)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(130))]
    flat = _flatten(ps)
    n = flat.numel()
    # each rank sees a different shard of the synthetic stream
    h, z, s, _ = next(synthetic_batches(2, 16, seed=rank))
    g = torch.zeros(n)
    off = 0
    for p in ps:
        k = p.numel()
        g[off:off + k] = (h.mean() + rank + 1) * torch.arange(k).float() / k  # stand-in local gradient
        off += _pad64(k)
    local = g.clone()
    dist.all_reduce(g)  # the trainer's only collective (sum); 1/world is folded into the optimizer
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(g, sum(gathered))
    gs = g / world
    norm = float(gs.norm())
    coef = min(1.0, 1.0 / (norm + 1e-6))
    m, v, ema = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    O.adamw_ema_step(flat, gs, m, v, ema, 1, 3e-4 * O.lr_lambda(0), clip_coef=coef, ema_first=True)
    allp = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    assert all(torch.equal(allp[0], a) for a in allp), 'replicas diverged'
    assert torch.equal(ps[2].data, flat[_pad64(37 * 11) + _pad64(5):][:130])  # params are views of the flat buffer
    # ---- the cached-latent reader under data parallelism: same window stream on every rank, rank r materialises
    #      batch k*world + r, all ranks take the same number of steps (else the all-reduce above would hang)
    from pathlib import Path
    from osu_dreamer_b200.data import DeviceFeeder, LatentWindows, batches
    sets = sorted(p for p in Path(cache).iterdir() if p.is_dir())
    mine = list(DeviceFeeder(LatentWindows(sets, 96, 4, -1, seed=7), 3, rank, world, device=None, depth=2))
    counts = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(mine)]))
    assert all(int(c) == len(mine) for c in counts) and len(mine) >= 2
    sig = torch.stack([b[0].double().sum() + 1000 * b[1].double().sum() for b in mine])  # one signature per batch
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    every = list(batches(LatentWindows(sets, 96, 4, -1, seed=7), 3, pin=False))  # the unsharded stream
    want = torch.stack([b[0].double().sum() + 1000 * b[1].double().sum() for b in every])
    for r in range(world):
        assert torch.equal(sigs[r], want[r::world][:len(mine)])
    if rank == 0:
        torch.save({'ok': True, 'steps_per_rank': len(mine), 'unsharded_batches': len(every)}, out)
    dist.destroy_process_group()


def test_ddp_host_logic_world2(tmp_path):
    import numpy as np
    rng = np.random.default_rng(0)
    cache = tmp_path / 'data'
    for ms in range(5):
        d = cache / f'set{ms}'
        d.mkdir(parents=True)
        l = 500 + 61 * ms
        np.save(d / 'h.npy', rng.standard_normal((128, l)).astype(np.float32))
        for k in range(3):
            np.savez(d / f'm{k}.latent.npz', z=rng.standard_normal((6, l)).astype(np.float32),
                     s=rng.standard_normal(32).astype(np.float32), labels=rng.random(5).astype(np.float32))
    out = str(tmp_path / 'r.pt')
    mp.spawn(_worker, args=(2, _free_port(), out, str(cache)), nprocs=2, join=True)
    r = torch.load(out)
    assert r['ok'] and r['steps_per_rank'] == r['unsharded_batches'] // 2
