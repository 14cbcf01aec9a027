"""Training data for `fit-denoiser`: the reference's cached-latent format and a synthetic source.

On-disk format (written by the reference's `encode-latents`, osu_dreamer/scripts/encode_latents.py:36-51):
`<data>/<mapset>/<map>.latent.npz` with arrays z [6,l], s [32], labels [5], and one `<data>/<mapset>/h.npy`
[128,l] per mapset.  `LatentBatch(h, z, s, labels)` mirrors osu_dreamer/data/modules/latent.py:21-25.
The reader keeps the reference's per-map window draw (random offset < seq_len, stride seq_len, at most `max_per_map`
windows, osu_dreamer/data/modules/latent.py:131-149) and its shuffle buffer (:112-129).  With `rng='global'` it
consumes the global torch / python generators exactly like the reference's single-process loader, so the two produce
the same sample stream (tests/test_host.py::test_latent_windows_match_reference_stream).

What is different is how bytes move, because one B200 consumes ~75 samples/s of 4.4 MB each at seq_len 8192:
* `h.npy` (one per mapset, shared by all its maps) is memory-mapped and kept open per mapset, and only the drawn
  windows are copied out of it -- the reference re-reads and converts the whole array for every map;
* `Prefetcher` collates and pins batches on a background thread (numpy / torch copies release the GIL), a few batches
  ahead of the training loop, instead of 13 worker processes that pickle every sample through a pipe.
tools/loader_bench.py times both on a synthetic cache.  The metric in bench.py uses the synthetic source.
"""
from __future__ import annotations

import os
import queue
import random
import threading
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
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
    """whole map, as osu_dreamer/data/modules/latent.py:78-84."""
    with np.load(f) as d:
        z, s, labels = (torch.from_numpy(d[k]).float() for k in ('z', 's', 'labels'))
    h = torch.from_numpy(np.load(f.parent / 'h.npy')).float()
    return LatentBatch(h, z, s, labels)


class _HCache:
    """memory-mapped `h.npy` per mapset directory (small LRU): windows are sliced out of the page cache."""

    def __init__(self, capacity: int = 4):
        self.capacity, self.maps = capacity, OrderedDict()

    def get(self, mapset: Path) -> np.ndarray:
        m = self.maps.get(mapset)
        if m is None:
            m = np.load(mapset / 'h.npy', mmap_mode='r')
            self.maps[mapset] = m
            while len(self.maps) > self.capacity:
                self.maps.popitem(last=False)
        else:
            self.maps.move_to_end(mapset)
        return m


def _window_f32(a: np.ndarray, i: int, n: int) -> Tensor:
    """a[..., i:i+n] as an owned, contiguous fp32 tensor (one copy; converts if the cache is not fp32)."""
    return torch.from_numpy(np.array(a[..., i:i + n], dtype=np.float32, order='C', copy=True))


class WindowRef(NamedTuple):
    """one training window, not yet copied: `h` is the mapset's memory-mapped array, `z` the map's latent."""
    h: np.ndarray
    z: np.ndarray
    s: Tensor
    labels: Tensor
    i: int
    n: int

    def load(self) -> LatentBatch:
        return LatentBatch(_window_f32(self.h, self.i, self.n), _window_f32(self.z, self.i, self.n), self.s, self.labels)


class LatentWindows:
    """rng='own': private generators seeded with `seed` (sorted file order, reproducible per epoch);
    rng='global': the reference's exact consumption of the global generators (latent.py:106-109,141-143).
    Iterating yields materialised `LatentBatch` samples; `refs()` yields the same stream as `WindowRef`s whose bytes
    are only touched when a batch is collated (so a rank can skip the batches of the other ranks for free)."""

    def __init__(self, mapsets, seq_len: int | None, shuffle_buffer_size: int = 1, max_per_map: int = -1, seed: int = 0,
                 rng: str = 'own'):
        if rng not in ('own', 'global'):
            raise ValueError("rng must be 'own' or 'global'")
        self.mapsets, self.seq_len = list(mapsets), seq_len
        self.buf = max(1, shuffle_buffer_size)
        self.max_per_map = max_per_map if max_per_map > 0 else 1 << 30
        self.mode = rng
        self.rng = random.Random(seed) if rng == 'own' else random
        self.gen = torch.Generator().manual_seed(seed) if rng == 'own' else None
        self.hcache = _HCache()

    def _windows(self, f: Path) -> Iterator[WindowRef]:
        with np.load(f) as d:
            z, s, labels = d['z'], torch.from_numpy(d['s']).float(), torch.from_numpy(d['labels']).float()
        h = self.hcache.get(f.parent)
        if self.seq_len is None:
            yield WindowRef(h, z, s, labels, 0, max(h.shape[-1], z.shape[-1]))
            return
        n = self.seq_len
        end = z.shape[-1] - n + 1
        if end < 1:
            return
        start = int(torch.randint(0, min(n, end), (), generator=self.gen))
        idx = torch.arange(start, end, n)
        idx = idx[torch.randperm(len(idx), generator=self.gen)[:min(self.max_per_map, len(idx))]]
        for i in idx.tolist():
            yield WindowRef(h, z, s, labels, i, n)

    def refs(self) -> Iterator[WindowRef]:
        if self.mode == 'global':
            random.seed(torch.initial_seed())  # latent.py:103-109 with no worker processes
            files = [f for m in self.mapsets for f in m.glob('*.latent.npz')]  # directory order, like the reference
        else:
            files = [f for m in self.mapsets for f in sorted(m.glob('*.latent.npz'))]
        stream = (w for f in files for w in self._windows(f))
        if self.mode == 'global' and self.buf <= 1:
            yield from stream
            return
        pool: list[WindowRef] = []
        for smp in stream:
            if len(pool) < self.buf:
                pool.append(smp)
                continue
            j = self.rng.randrange(len(pool))
            yield pool[j]
            pool[j] = smp
        self.rng.shuffle(pool)
        yield from pool

    def __iter__(self) -> Iterator[LatentBatch]:
        return (r.load() for r in self.refs())


def _alloc_batch(B: int, A: int, E: int, n: int, S: int, pin: bool):
    pin = pin and torch.cuda.is_available()
    return (torch.empty(B, A, n, dtype=torch.float32, pin_memory=pin), torch.empty(B, E, n, dtype=torch.float32, pin_memory=pin),
            torch.empty(B, S, dtype=torch.float32, pin_memory=pin), torch.empty(B, 5, dtype=torch.float32, pin_memory=pin))


_COPY_THREADS = max(1, min(8, int(os.environ.get('OSD_LOADER_THREADS', '4'))))
_pool: ThreadPoolExecutor | None = None


def _copy_pool() -> ThreadPoolExecutor:
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(_COPY_THREADS, thread_name_prefix='osd-loader')
    return _pool


def _collate(refs: list, pin: bool, out=None):
    """[B, ...] host batch written once: every window goes straight from the page cache into its slot of `out`
    (a recycled -- typically pinned -- buffer set from _alloc_batch) or of a fresh allocation."""
    B, n = len(refs), refs[0].n
    if out is None:
        out = _alloc_batch(B, refs[0].h.shape[0], refs[0].z.shape[0], n, refs[0].s.shape[0], pin)
    h, z, s, labels = out
    hn, zn = h.numpy(), z.numpy()

    def one(j):
        r = refs[j]
        np.copyto(hn[j], r.h[..., r.i:r.i + n], casting='same_kind')  # releases the GIL: windows copy in parallel
        np.copyto(zn[j], r.z[..., r.i:r.i + n], casting='same_kind')
        s[j].copy_(r.s)
        labels[j].copy_(r.labels)

    if _COPY_THREADS > 1 and B > 1 and hn[0].nbytes >= (1 << 20):
        list(_copy_pool().map(one, range(B)))
    else:
        for j in range(B):
            one(j)
    return out


def batch_refs(windows, batch_size: int, rank: int = 0, world: int = 1):
    """groups of `batch_size` windows for this rank; drop_last.  Data parallel: every rank walks the same window
    stream (same seeds -> same order, no bytes touched) and keeps group k only if k % world == rank; a trailing run of
    fewer than `world` groups is dropped so that all ranks take the same number of steps."""
    cur, n, mine = [], 0, None
    for w in windows.refs() if hasattr(windows, 'refs') else windows:
        cur.append(w)
        if len(cur) == batch_size:
            if n % world == rank:
                mine = cur
            if n % world == world - 1:  # the round is complete: every rank has its group
                yield mine
                mine = None
            cur, n = [], n + 1


def batches(windows, batch_size: int, rank: int = 0, world: int = 1, pin: bool = True):
    """collated [B, ...] (pinned) host batches of this rank, freshly allocated (see DeviceFeeder for the recycled,
    overlapped path the training loop uses)."""
    for grp in batch_refs(windows, batch_size, rank, world):
        if isinstance(grp[0], WindowRef):
            yield _collate(grp, pin)
        else:
            out = tuple(torch.stack(x) for x in zip(*grp))
            yield tuple(t.pin_memory() for t in out) if pin and torch.cuda.is_available() else out


class Prefetcher:
    """Runs a batch iterator on a background thread, `depth` batches ahead (collation, pinning and the page-cache
    reads overlap the GPU step).  Exceptions of the producer re-raise in the consumer; `close()` stops it early."""

    _END = object()

    def __init__(self, it, depth: int = 3):
        self.q: queue.Queue = queue.Queue(maxsize=max(1, depth))
        self.stop = threading.Event()
        self.t = threading.Thread(target=self._run, args=(iter(it),), daemon=True)
        self.t.start()

    def _put(self, item) -> bool:
        while not self.stop.is_set():
            try:
                self.q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, it):
        try:
            for b in it:
                if not self._put(b):
                    return
            self._put(self._END)
        except BaseException as e:  # noqa: BLE001 -- handed to the consumer
            self._put(e)

    def __iter__(self):
        return self

    def __next__(self):
        item = self.q.get()
        if item is self._END:
            self.t.join(timeout=5)
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item

    def close(self):
        self.stop.set()
        self.t.join(timeout=5)


class DeviceFeeder:
    """Background thread: window groups -> a ring of recycled pinned host buffers -> asynchronous H2D copies on a
    side stream, `depth` batches ahead of the consumer.  Iterating yields DEVICE batches; the consumer's current stream
    is made to wait on the copy's event, so the copy of batch k+1.. overlaps the compute of batch k and nothing
    synchronises the host.  A ring slot is refilled only after its copy event has completed.
    With device=None (no GPU) it degrades to recycled host buffers handed out as clones."""

    _END = object()

    def __init__(self, windows, batch_size: int, rank: int = 0, world: int = 1, device=None, depth: int = 3):
        self.device = torch.device(device) if device is not None else None
        self.cuda = self.device is not None and self.device.type == 'cuda'
        self.depth = max(1, depth)
        self.q: queue.Queue = queue.Queue(maxsize=self.depth)
        self.stop = threading.Event()
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.t = threading.Thread(target=self._run, args=(batch_refs(windows, batch_size, rank, world),), daemon=True)
        self.t.start()

    def _put(self, item) -> bool:
        while not self.stop.is_set():
            try:
                self.q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, groups):
        try:
            ring, k = [], 0
            slots = self.depth + 2  # being filled + queued + in the consumer's hands
            for grp in groups:
                if len(ring) < slots:
                    r0 = grp[0]
                    ring.append([_alloc_batch(len(grp), r0.h.shape[0], r0.z.shape[0], r0.n, r0.s.shape[0], self.cuda), None])
                slot = ring[k % slots]
                k += 1
                if slot[1] is not None:
                    slot[1].synchronize()  # the copy that last read this slot has finished
                if slot[0][0].shape[0] != len(grp) or slot[0][0].shape[-1] != grp[0].n:
                    slot[0] = _alloc_batch(len(grp), grp[0].h.shape[0], grp[0].z.shape[0], grp[0].n, grp[0].s.shape[0], self.cuda)
                host = _collate(grp, self.cuda, out=slot[0])
                if self.cuda:
                    with torch.cuda.stream(self.stream):
                        dev = tuple(t.to(self.device, non_blocking=True) for t in host)
                        ev = torch.cuda.Event()
                        ev.record(self.stream)
                    slot[1] = ev
                    item = (dev, ev)
                else:
                    item = (tuple(t.clone() for t in host), None)
                if not self._put(item):
                    return
            self._put(self._END)
        except BaseException as e:  # noqa: BLE001 -- handed to the consumer
            self._put(e)

    def __iter__(self):
        return self

    def __next__(self):
        item = self.q.get()
        if item is self._END:
            self.t.join(timeout=5)
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        dev, ev = item
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in dev:
                t.record_stream(cur)
        return dev

    def close(self):
        self.stop.set()
        self.t.join(timeout=5)


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
