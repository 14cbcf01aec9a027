"""Throughput of the cached-latent reader (CPU only): samples/s and GB/s of collated [B, ...] batches at the
training shape, for (a) the reference's access pattern -- whole `h.npy` re-read and converted for every map
(osu_dreamer/data/modules/latent.py:78-84,131-149; the unmodified reference class when importable, else a restatement)
and (b) this package's reader (memory-mapped h, window copies only; DeviceFeeder = recycled pinned ring + background
thread + side-stream H2D when a GPU is present).
Usage: python tools/loader_bench.py [cache_dir] [--sets N] [--maps M] [--frames L] [--seq-len S] [--batch B]"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from osu_dreamer_b200.data import DeviceFeeder, LatentWindows, batches, load_latents, LatentBatch

ap = argparse.ArgumentParser()
ap.add_argument('cache', nargs='?')
ap.add_argument('--sets', type=int, default=24)
ap.add_argument('--maps', type=int, default=4)
ap.add_argument('--frames', type=int, default=20000)
ap.add_argument('--seq-len', type=int, default=8192)
ap.add_argument('--batch', type=int, default=16)
args = ap.parse_args()

tmp = None
root = Path(args.cache) if args.cache else Path((tmp := tempfile.TemporaryDirectory()).name)
if not any(root.iterdir()) if root.exists() else True:
    root.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(0)
    for ms in range(args.sets):
        d = root / f'set{ms:03d}'
        d.mkdir()
        l = args.frames + 97 * ms
        np.save(d / 'h.npy', rng.standard_normal((128, l), dtype=np.float32))
        for k in range(args.maps):
            np.savez(d / f'm{k}.latent.npz', z=rng.standard_normal((6, l), dtype=np.float32),
                     s=rng.standard_normal(32, dtype=np.float32), labels=rng.random(5, dtype=np.float32))
sets = sorted(p for p in root.iterdir() if p.is_dir())


def reference_stream():
    try:
        from oracle import refimport
        if refimport.available():
            refimport._install_stubs()
            if refimport.REFERENCE_ROOT not in sys.path:
                sys.path.insert(0, refimport.REFERENCE_ROOT)
            from osu_dreamer.data.modules.latent import LatentDataset
            return 'reference LatentDataset (single process)', LatentDataset(sets, args.seq_len, 1, -1)
    except Exception:
        pass

    def port():
        for m in sets:
            for f in sorted(m.glob('*.latent.npz')):
                h, z, s, lab = load_latents(f)
                end = z.size(-1) - args.seq_len + 1
                for i in range(0, max(0, end), args.seq_len):
                    yield LatentBatch(h[..., i:i + args.seq_len].clone(), z[..., i:i + args.seq_len].clone(), s, lab)
    return 'restatement of the reference access pattern', port()


def run(name, it):
    t0 = time.perf_counter()
    n = nbytes = 0
    for b in it:
        n += b[0].shape[0]
        nbytes += sum(t.numel() * t.element_size() for t in b)
    dt = time.perf_counter() - t0
    row = {'reader': name, 'samples': n, 'seconds': round(dt, 3), 'samples_per_s': round(n / dt, 1), 'GB_per_s': round(nbytes / dt / 1e9, 2)}
    print(json.dumps(row), flush=True)
    return row

torch.manual_seed(0)
rows = []
for rep in range(2):  # second round: page cache warm for everyone
    name, ref = reference_stream()
    rows.append(run(name, batches(ref, args.batch, pin=False)))
    rows.append(run('b200 reader (mmap h, window copies)', batches(LatentWindows(sets, args.seq_len, 1, -1, seed=rep), args.batch, pin=False)))
    dev = 'cuda' if torch.cuda.is_available() else None
    rows.append(run(f'b200 DeviceFeeder (recycled {"pinned " if dev else ""}ring, background thread, device={dev})', DeviceFeeder(LatentWindows(sets, args.seq_len, 1, -1, seed=rep), args.batch, device=dev)))
os.makedirs('gpurun_out', exist_ok=True)
json.dump({'cache': {'mapsets': len(sets), 'maps_per_set': args.maps, 'frames': args.frames, 'seq_len': args.seq_len, 'batch': args.batch},
           'cpu_count': os.cpu_count(), 'rows': rows}, open('gpurun_out/loader_bench.json', 'w'), indent=1)
