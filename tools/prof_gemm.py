"""Small driver for ncu: the layer's proj_vg (K = 512) and dgrad proj_vg (K = 2816) GEMMs at T = 131072 tokens, twice each."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib
T, bf = 131072, torch.bfloat16
r = lambda *s: torch.randn(*s, device='cuda').to(bf)
A, B, C = r(T, 512), r(2816, 512), torch.empty(T, 2816, dtype=bf, device='cuda')
A2, B2, C2 = r(T, 2816), r(2816, 512), torch.empty(T, 512, dtype=bf, device='cuda')
bias = torch.randn(2816, device='cuda')
for _ in range(2):
    lib.gemm(A, B, C, bias=bias)
    lib.gemm(A2, B2, C2, b_major=lib.MAJOR_MN)
torch.cuda.synchronize()
print('done')
