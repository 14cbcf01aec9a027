"""ncu driver: forward attention variants at the bench shape, one profiled launch each (after a warm-up round).
    ncu --set full --import-source on --profile-from-start off -o gpurun_out/<tag>_fwd_variants python tools/prof_fwd_variants.py 7 9 10 11"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from osu_dreamer_b200 import lib

variants = [int(a) for a in sys.argv[1:]] or [7, 9]
B, L = 16, 8192
qkv = torch.randn(B * L, 3072, device='cuda').to(torch.bfloat16)
bound = torch.tensor([14.0], device='cuda')
for v in variants:
    lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=v)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for v in variants:
    lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=v)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
