"""Training data for `fit-denoiser`: the reference's cached-latent format and a synthetic source.

On-disk format (written by the reference's `encode-latents`, osu_dreamer/scripts/encode_latents.py:36-51):
`<data>/<mapset>/<map>.latent.npz` with arrays z [6,l], s [32], labels [5], and one `<data>/<mapset>/h.npy`
[128,l] per mapset.  `LatentBatch(h, z, s, labels)` mirrors osu_dreamer/data/modules/latent.py:21-25.
This reader is a single-process windowing sampler (the reference uses 13 DataLoader workers); it keeps the
reference's per-map window draw (random offset < seq_len, stride seq_len, at most `max_per_map` windows) and
shuffle buffer.  The metric in bench.py uses the synthetic source.
"""
from __future__ import annotations

import random
from pathlib import Path
from typing import Iterator, NamedTuple

import numpy as np
import torch
from torch import Tensor


class LatentBatch(NamedTuple):
    h: Tensor       # audio features (chunk rate)   [A, l]
    z: Tensor       # chart latent                  [E, l]
    s: Tensor       # per-map style code            [S]
    labels: Tensor  # difficulty labels             [5]


def split_mapsets(data_dir: Path, pattern: str, max_val_count: int, max_val_frac: float):
    """whole-mapset hold-out, same rule as osu_dreamer/data/modules/beatmap.py:33-71."""
    if not data_dir.exists():
        raise ValueError(f'data dir `{data_dir}` does not exist, generate dataset first')
    counts = [(d, sum(1 for _ in d.glob(pattern))) for d in data_dir.iterdir() if d.is_dir()]
    full = sum(c for _, c in counts)
    if full == 0:
        raise ValueError(f'data dir `{data_dir}` is empty, generate dataset first')
    limit = min(max_val_count, int(full * max_val_frac))
    if not (0 < limit < full):
        raise ValueError(f'invalid validation size {limit} for {full} maps')
    train, val, nval = [], [], 0
    for d, c in counts:
        if nval + c > limit:
            train.append(d)
        else:
            val.append(d)
            nval += c
    return train, val


def load_latents(f: Path) -> LatentBatch:
    with np.load(f) as d:
        z, s, labels = (torch.from_numpy(d[k]).float() for k in ('z', 's', 'labels'))
    h = torch.from_numpy(np.load(f.parent / 'h.npy')).float()
    return LatentBatch(h, z, s, labels)


class LatentWindows:
    def __init__(self, mapsets, seq_len: int | None, shuffle_buffer_size: int = 1, max_per_map: int = -1, seed: int = 0):
        self.mapsets, self.seq_len = list(mapsets), seq_len
        self.buf = max(1, shuffle_buffer_size)
        self.max_per_map = max_per_map if max_per_map > 0 else 1 << 30
        self.rng = random.Random(seed)
        self.gen = torch.Generator().manual_seed(seed)

    def _windows(self, f: Path) -> Iterator[LatentBatch]:
        h, z, s, labels = load_latents(f)
        if self.seq_len is None:
            yield LatentBatch(h, z, s, labels)
            return
        end = z.size(-1) - self.seq_len + 1
        if end < 1:
            return
        start = int(torch.randint(0, min(self.seq_len, end), (), generator=self.gen))
        idx = torch.arange(start, end, self.seq_len)
        idx = idx[torch.randperm(len(idx), generator=self.gen)[:min(self.max_per_map, len(idx))]]
        for i in idx.tolist():
            yield LatentBatch(h[..., i:i + self.seq_len].clone(), z[..., i:i + self.seq_len].clone(), s, labels)

    def __iter__(self) -> Iterator[LatentBatch]:
        files = [f for m in self.mapsets for f in sorted(m.glob('*.latent.npz'))]
        stream = (w for f in files for w in self._windows(f))
        pool: list[LatentBatch] = []
        for smp in stream:
            if len(pool) < self.buf:
                pool.append(smp)
                continue
            j = self.rng.randrange(len(pool))
            yield pool[j]
            pool[j] = smp
        self.rng.shuffle(pool)
        yield from pool


def batches(windows, batch_size: int, rank: int = 0, world: int = 1, pin: bool = True):
    """collate windows into [B, ...] pinned host batches; rank r takes every world-th batch (drop_last)."""
    cur, n = [], 0
    for w in windows:
        cur.append(w)
        if len(cur) == batch_size:
            if n % world == rank:
                out = tuple(torch.stack(x) for x in zip(*cur))
                yield tuple(t.pin_memory() for t in out) if pin and torch.cuda.is_available() else out
            cur, n = [], n + 1


def synthetic_batches(batch_size: int, seq_len: int, seed: int = 0, a_dim: int = 128, emb_dim: int = 6, style_dim: int = 32):
    """endless random batches with the statistics of real latents: per-frame RMS-normalised z
    (osu_dreamer/models/latent/model.py:62-65) and RMS-normalised s (:55-59)."""
    g = torch.Generator().manual_seed(seed)
    while True:
        h = torch.randn(batch_size, a_dim, seq_len, generator=g)
        z = torch.randn(batch_size, emb_dim, seq_len, generator=g)
        z = z * z.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()
        s = torch.randn(batch_size, style_dim, generator=g)
        s = s * s.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()
        yield (h, z, s, 10 * torch.rand(batch_size, 5, generator=g))
