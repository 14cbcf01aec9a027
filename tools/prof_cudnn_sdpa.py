"""Driver for an ncu capture of the library kernel to beat: cuDNN's scaled_dot_product_attention forward + backward at the
layer's shape (q, k, v [16,16,8192,64] bf16, contiguous), next to this repo's forward (variant 7) and single-pass backward.
    ncu --set full --clock-control none -o gpurun_out/<tag>_cudnn_sdpa python tools/prof_cudnn_sdpa.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from osu_dreamer_b200 import lib

B, H, L, d = 16, 16, 8192, 64
g = torch.Generator(device='cuda').manual_seed(0)
q, k, v = (torch.randn(B, H, L, d, device='cuda', generator=g).to(torch.bfloat16).requires_grad_(True) for _ in range(3))
do = torch.randn(B, H, L, d, device='cuda', generator=g).to(torch.bfloat16)
qkv = torch.randn(B * L, 3072, device='cuda', generator=g).to(torch.bfloat16)
dy = torch.randn(B * L, 1024, device='cuda', generator=g).to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
    o = F.scaled_dot_product_attention(q, k, v)
    o.backward(do)
y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=7)
dqkv = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
