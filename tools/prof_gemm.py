"""Small driver for ncu.  Default: the layer's proj_vg (K = 512) and dgrad proj_vg (K = 2816) GEMMs at T = 131072 tokens;
'wgrad': the qkv and out_proj weight-gradient GEMMs (MN-major operands, split-K reduce-add).  Twice each."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib
T, bf = 131072, torch.bfloat16
r = lambda *s: torch.randn(*s, device='cuda').to(bf)
if len(sys.argv) > 1 and sys.argv[1] == 'wgrad':
    jobs = []
    for M, N in ((3072, 512), (512, 1024)):
        A, B, C = r(T, M), r(T, N), torch.zeros(M, N, device='cuda')
        sk = lib.gemm_split_k(M, N, T)
        jobs.append(lambda A=A, B=B, C=C, sk=sk: lib.gemm(A, B, C, a_major=lib.MAJOR_MN, b_major=lib.MAJOR_MN, epi=lib.EPI_ATOMIC, split_k=sk))
else:
    A, B, C = r(T, 512), r(2816, 512), torch.empty(T, 2816, dtype=bf, device='cuda')
    A2, B2, C2 = r(T, 2816), r(2816, 512), torch.empty(T, 512, dtype=bf, device='cuda')
    bias = torch.randn(2816, device='cuda')
    jobs = [lambda: lib.gemm(A, B, C, bias=bias), lambda: lib.gemm(A2, B2, C2, b_major=lib.MAJOR_MN)]
for _ in range(2):
    for j in jobs:
        j()
torch.cuda.synchronize()
print('done')
