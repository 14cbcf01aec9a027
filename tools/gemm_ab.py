"""GEMM shapes of one denoiser layer at the bench size (T = 16 x 8192 tokens): time per launch and TFLOP/s.
Run twice for an A/B: OSD_GEMM_EW=4 forces the 4-epilogue-warp layout, OSD_GEMM_PAIR=0 / 1 switches the CTA-pair kernel off /
on for every eligible shape (default: K >= 1024)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib

T = 131072
dev = 'cuda'
bf, f32 = torch.bfloat16, torch.float32


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rnd(*s, dtype=bf):
    return torch.randn(*s, device=dev).to(dtype)


rows = []
# forward (A, B K-major)
for name, N, K, cdt in (('proj_cl', 512, 128, bf), ('out_proj', 512, 1024, f32), ('proj_vg', 2816, 512, bf), ('proj_o', 512, 1408, f32)):
    A, Bm, C = rnd(T, K), rnd(N, K), torch.empty(T, N, dtype=cdt, device=dev)
    bias = rnd(N, dtype=f32)
    ms = timeit(lambda: lib.gemm(A, Bm, C, bias=bias))
    rows.append((name, N, K, ms))
x, w, b = rnd(T, 512), rnd(3072, 512), rnd(3072, dtype=f32)
qw, kw = rnd(64, dtype=f32), rnd(64, dtype=f32)
rope = lib.rope_table(8192, dev)
raw = torch.empty(T, 3072, dtype=bf, device=dev)
rows.append(('qkv_proj+norm+rope (train: raw copy)', 3072, 512, timeit(lambda: lib.qkv_proj(x, w, b, qw, kw, rope, 8192, raw_out=raw))))
rows.append(('qkv_proj+norm+rope (inference)', 3072, 512, timeit(lambda: lib.qkv_proj(x, w, b, qw, kw, rope, 8192))))
# dgrad: dX[T,N] = dY[T,K] W[K,N] (B MN-major)
for name, N, K in (('dgrad proj_o (dhn)', 1408, 512), ('dgrad proj_vg (dz2)', 512, 2816), ('dgrad out_proj (dy)', 1024, 512), ('dgrad qkv (dz)', 512, 3072)):
    A, Bm, C = rnd(T, K), rnd(K, N), torch.empty(T, N, dtype=bf, device=dev)
    ms = timeit(lambda: lib.gemm(A, Bm, C, b_major=lib.MAJOR_MN))
    rows.append((name, N, K, ms))
# wgrad: dW[M,N] += dY[T,M]^T X[T,N] (both MN-major, split-K reduce-add into fp32)
for name, M, N in (('wgrad qkv', 3072, 512), ('wgrad out_proj', 512, 1024), ('wgrad proj_vg', 2816, 512), ('wgrad proj_o', 512, 1408)):
    A, Bm, C = rnd(T, M), rnd(T, N), torch.zeros(M, N, dtype=f32, device=dev)
    sk = lib.gemm_split_k(M, N, T)
    ms = timeit(lambda: lib.gemm(A, Bm, C, a_major=lib.MAJOR_MN, b_major=lib.MAJOR_MN, epi=lib.EPI_ATOMIC, split_k=sk))
    rows.append((f'{name} [{M}x{N}] split {sk}', N, M, ms))  # the TF/s line below uses 2 T N K with K := M
for name, N, K, ms in rows:
    print(f'{name:40s} N={N:5d} K={K:5d}  {ms * 1e3:8.1f} us  {2 * T * N * K / ms / 1e9:7.0f} TF/s', flush=True)
print('sum', sum(r[3] for r in rows))
