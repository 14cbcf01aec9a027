"""Small driver for ncu: the fused QKV projection (bias + per-head RMSNorm + RoPE epilogue, raw copy for training) at
T = 16 x 8192 tokens, twice."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib
T, bf, f32 = 131072, torch.bfloat16, torch.float32
r = lambda *s, dtype=bf: torch.randn(*s, device='cuda').to(dtype)
x, w, b = r(T, 512), r(3072, 512), r(3072, dtype=f32)
qw, kw = r(64, dtype=f32), r(64, dtype=f32)
rope = lib.rope_table(8192, 'cuda')
raw = torch.empty(T, 3072, dtype=bf, device='cuda')
for _ in range(2):
    lib.qkv_proj(x, w, b, qw, kw, rope, 8192, raw_out=raw)
torch.cuda.synchronize()
print('done')
