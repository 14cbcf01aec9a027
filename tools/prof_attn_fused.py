"""ncu driver: one launch of the single-pass attention backward (and the two-pass kernels for comparison)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
dy = torch.randn(B * L, 1024, device='cuda').to(torch.bfloat16)
y, lse = lib.attn_fwd(qkv, B, L)
for _ in range(2):
    d = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
torch.cuda.synchronize()
print('done')
