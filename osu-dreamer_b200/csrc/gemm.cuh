// tcgen05 GEMM: C[M,N] (+)= A[M,K] * B[N,K]^T, fp32 accumulation in TMEM, operands staged by TMA.
#pragma once
#include "common.h"

namespace osd {

enum { MAJOR_K = 0, MAJOR_MN = 1 };
enum { ELEM_BF16 = 0, ELEM_TF32 = 1 };
enum {
  EPI_STORE = 0,   // C = acc + bias
  EPI_SILU = 1,    // C = silu(acc + bias)
  EPI_ATOMIC = 2,  // C += acc  (fp32 red.global.add; split-K / gradient accumulation)
  EPI_QKV = 3,     // bias + per-head RMSNorm(q,k) + RoPE (common/attn.py:75-81 of the reference)
};

struct GemmArgs {
  const void* A = nullptr;  // MAJOR_K : [M, K] row-major, leading dim lda (elements)
  const void* B = nullptr;  // MAJOR_MN: [K, M] row-major, leading dim lda   (same for B with N)
  int a_major = MAJOR_K, b_major = MAJOR_K;
  int64_t lda = 0, ldb = 0;
  int M = 0, N = 0, K = 0;
  int elem = ELEM_BF16;
  int epi = EPI_STORE;
  void* C = nullptr;
  int64_t ldc = 0;
  int c_fp32 = 0;
  const float* bias = nullptr;
  int split_k = 1;
  // 3x split precision ("fp32-grade" path): A is [M, 2K] = (hi | lo), B is [N, 2K] = (hi | lo), both K-major;
  // the kernel runs the three partial products A_hi B_hi + A_lo B_hi + A_hi B_lo as one 3K-long reduction.
  // bf16 elements: hi = bf16(x), lo = bf16(x - hi), ~1e-5 of the fp32 product; tf32 elements (fp32 storage): hi = tf32(x),
  // lo = x - hi (the tensor core reads its top 19 bits), ~1e-6 -- the latent blocks (latent_tc.cu).
  int split3 = 0;
  // bf16 outputs written as (hi | lo) pairs: C is [M, 2N], hi in columns [0,N), lo = bf16(x - hi) in [N,2N)
  int c_split = 0;
  // EPI_QKV
  const float* qnorm_w = nullptr;  // [64]
  const float* knorm_w = nullptr;  // [64]
  const float* rope = nullptr;     // [L][2][32] fp32 (cos | sin)
  int L = 0;                       // sequence length (row m is position m % L)
  int dh = 1024;                   // n_heads * head_dim: columns [0,dh) q, [dh,2dh) k, [2dh,3dh) v
  void* raw_out = nullptr;         // optional bf16 [M, N] (ld = ldc): pre-norm q,k,v (saved for backward)
};

int launch_gemm(const GemmArgs& a, cudaStream_t stream);
// split-K factor for a reduce-add GEMM (weight gradients: small M x N, K = tokens): the split whose work items fill whole
// waves of the persistent grid (CTAs, or CTA pairs where launch_gemm will use them) with the fewest k-blocks per wave
int gemm_split_for(int M, int N, int K);

}  // namespace osd
