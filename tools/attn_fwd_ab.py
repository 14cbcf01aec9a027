"""attention forward: value check against torch + timing at the bench / sampling shapes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib


def timeit(fn, iters=5, warm=2):
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


B, L = 2, 1000
g = torch.Generator().manual_seed(1)
qkv = torch.randn(B * L, 3072, generator=g).cuda().to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
for variant, bl in ((7, bound), (7, None), (17, bound), (17, None)):
    y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bl, variant=variant)
    q, k, v = qkv.float().view(B, L, 3, 16, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, 1024)
    lref = torch.logsumexp(s, -1)
    print('variant', variant, 'fixed' if bl is not None else 'online', 'y err', float((y.float() - ref).abs().max() / ref.abs().max()),
          'lse err', float((lse - lref).abs().max()), flush=True)
import json
res = []
for B, L in ((16, 8192), (32, 8192), (8, 2048), (8, 4096), (8, 16384), (4, 1500)):
    qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
    fl = 4.0 * B * 16 * L * L * 64
    for variant in (7, 17, 7, 17, 7, 17):
        t = timeit(lambda: lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=variant), iters=10, warm=3)
        print(f'B={B} L={L} variant {variant}: fwd {t:.3f} ms ({fl / t / 1e9:.0f} TF/s)', flush=True)
        res.append({'B': B, 'L': L, 'variant': variant, 'ms': t, 'tflops': fl / t / 1e9})
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'attn_fwd_ab.json'), 'w'), indent=1)
