"""Forward attention, variant 7 (one q tile per CTA, two CTAs per SM) against variant 18 (two q tiles per CTA, sixteen softmax
warps) over the sequence length at a constant token count: where `attn_fwd_variant(L)` (csrc/model.cu) should switch."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib

rows = []
for L in (1024, 1482, 2048, 3072, 4096, 5120, 6144, 8192):
    B = max(1, 131072 // L)
    qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
    bound = torch.tensor([14.0], device='cuda')
    r = {'L': L, 'B': B}
    for v in (7, 18, 7, 18):
        for _ in range(3):
            lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=v)
        e1.record()
        torch.cuda.synchronize()
        r[f'v{v}_ms'] = min(r.get(f'v{v}_ms', 1e9), e0.elapsed_time(e1) / 10)
    r['v18_over_v7'] = r['v18_ms'] / r['v7_ms']
    rows.append(r)
    print(json.dumps(r), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/attn_fwd_threshold.json', 'w'), indent=1)
