"""A/B of the attention backward kernels on the GPU: single-pass (attn_bwd_fused.cu) against the two-kernel
path (attn_bwd.cu, parity-tested against autograd) -- values, then timing at the bench shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))


def check(B, L, spread=False):
    g = torch.Generator().manual_seed(L + B)
    qkv = torch.randn(B * L, 3072, generator=g)
    if spread:  # one query row with huge scores: its lse is > 96 octaves above its neighbours -> device-side fallback
        qkv[3, :1024] *= 40.0
    qkv = qkv.cuda().to(torch.bfloat16)
    dy = torch.randn(B * L, 1024, generator=g).cuda().to(torch.bfloat16)
    y, lse = lib.attn_fwd(qkv, B, L)
    ref = lib.attn_bwd(qkv, y, dy, lse, B, L)
    got = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
    torch.cuda.synchronize()
    out = {n: rel(got[:, s], ref[:, s]) for n, s in (('dq', slice(0, 1024)), ('dk', slice(1024, 2048)), ('dv', slice(2048, 3072)))}
    print(f'B={B} L={L} spread={spread}', {k: f'{v:.2e}' for k, v in out.items()}, 'finite', bool(torch.isfinite(got.float()).all()), flush=True)
    return max(out.values())


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == '__main__':
    worst = 0.0
    for B, L in [(1, 128), (1, 256), (2, 320), (1, 1000), (2, 2048)]:
        worst = max(worst, check(B, L))
    worst = max(worst, check(1, 512, spread=True))
    print('worst', worst, flush=True)
    for B, L in [(8, 8192), (16, 8192)]:
        qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
        dy = torch.randn(B * L, 1024, device='cuda').to(torch.bfloat16)
        y, lse = lib.attn_fwd(qkv, B, L)
        fl = 8.0 * B * 16 * L * L * 64  # 2 x forward
        t2 = timeit(lambda: lib.attn_bwd(qkv, y, dy, lse, B, L))
        t1 = timeit(lambda: lib.attn_bwd_fused(qkv, y, dy, lse, B, L))
        print(f'B={B} L={L}: two-pass {t2:.3f} ms ({fl / t2 / 1e9:.0f} TF/s)  fused {t1:.3f} ms ({fl / t1 / 1e9:.0f} TF/s)', flush=True)
