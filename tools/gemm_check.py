"""osd_gemm against torch.matmul over operand layouts (K-major / MN-major A and B), outputs (bf16 / fp32 / split-K reduce-add),
ragged M and the fused QKV epilogue (checked against the non-pair kernel, which the model parity tests cover).
OSD_GEMM_PAIR=1 in the environment routes the wide bf16 shapes to the cta_group::2 kernel.  Prints one line per case and
exits non-zero on a mismatch."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from osu_dreamer_b200 import lib

torch.backends.cuda.matmul.allow_tf32 = False
dev, bf, f32 = 'cuda', torch.bfloat16, torch.float32
g = torch.Generator(device=dev).manual_seed(0)
bad = 0


def rnd(*s, dtype=bf):
    return torch.randn(*s, device=dev, generator=g).to(dtype)


def check(name, got, ref, tol):
    global bad
    err = float((got.float() - ref).abs().max() / ref.abs().max())
    ok = err <= tol and bool(torch.isfinite(got.float()).all())
    bad += 0 if ok else 1
    print(f'{"ok  " if ok else "FAIL"} {name:60s} err {err:.2e}', flush=True)


for M in (32768, 20000 + 77 * 8, 512 * 37):
    for N, K in ((512, 1024), (2816, 512), (1024, 512), (512, 1408), (1408, 512), (768, 64), (96, 512)):
        A, B = rnd(M, K), rnd(N, K)
        ref = A.float() @ B.float().T
        bias = rnd(N, dtype=f32)
        check(f'K-major bf16 out + bias M={M} N={N} K={K}', lib.gemm(A, B, torch.empty(M, N, dtype=bf, device=dev), bias=bias), ref + bias, 6e-3)
        check(f'K-major fp32 out M={M} N={N} K={K}', lib.gemm(A, B, torch.empty(M, N, dtype=f32, device=dev)), ref, 1e-5)
        Bt = B.T.contiguous()  # [K, N]
        check(f'B MN-major bf16 out M={M} N={N} K={K}', lib.gemm(A, Bt, torch.empty(M, N, dtype=bf, device=dev), b_major=lib.MAJOR_MN), ref, 6e-3)
# wgrad shape: A [K=T, M] MN-major, B [K=T, N] MN-major, split-K reduce-add into fp32
T = 16384
for M, N in ((3072, 512), (512, 1024), (2816, 512), (512, 1408)):
    A, B = rnd(T, M), rnd(T, N)
    ref = A.float().T @ B.float()
    C = torch.zeros(M, N, dtype=f32, device=dev)
    for sk in (16, lib.gemm_split_k(M, N, T)):
        C.zero_()
        lib.gemm(A, B, C, a_major=lib.MAJOR_MN, b_major=lib.MAJOR_MN, epi=lib.EPI_ATOMIC, split_k=sk)
        check(f'wgrad MN/MN split-K {sk} M={M} N={N} K={T}', C, ref, 2e-5)
# fused QKV epilogue: bitwise against the single-CTA kernel (subprocess-free: the pair switch is per process, so compare
# against a torch restatement at bf16 tolerance instead)
L, Bn = 4096, 6
x, w, b = rnd(Bn * L, 512), rnd(3072, 512) * 0.05, rnd(3072, dtype=f32) * 0.1
qw, kw = 1 + 0.1 * rnd(64, dtype=f32), 1 + 0.1 * rnd(64, dtype=f32)
rope = lib.rope_table(L, dev)
raw = torch.empty(Bn * L, 3072, dtype=bf, device=dev)
out = lib.qkv_proj(x, w, b, qw, kw, rope, L, raw_out=raw)
z = x.float() @ w.float().T + b
check('qkv raw projections', raw, z, 6e-3)
zq = z.view(Bn, L, 3, 16, 64)
eps = torch.finfo(torch.float32).eps
inv_freq = 10000 ** (torch.arange(0, 64, 2, device=dev).float() / -64)
ang = torch.arange(L, device=dev).float()[:, None] * inv_freq[None]
cos, sin = ang.cos()[None, :, None, :], ang.sin()[None, :, None, :]
res = []
for i, wn in ((0, qw), (1, kw)):
    t = zq[:, :, i]
    t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps) * wn
    a_, b_ = t[..., :32], t[..., 32:]
    res.append(torch.cat([a_ * cos - b_ * sin, a_ * sin + b_ * cos], -1))
res.append(zq[:, :, 2])
ref = torch.stack(res, 2).reshape(Bn * L, 3072)
check('qkv norm + rope epilogue', out, ref, 1e-2)
print('mode', 'pair' if os.environ.get('OSD_GEMM_PAIR') == '1' else 'single', 'failures', bad)
sys.exit(1 if bad else 0)
