"""Small driver for ncu: two rounds of the attention kernels the model runs (forward variant 18, single-pass
backward = tilemax + stats + fused + dq convert + the two gated fallback launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
dy = torch.randn(B * L, 1024, device='cuda').to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
for _ in range(2):
    y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=18)
    dqkv = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
torch.cuda.synchronize()
print('done')
