// Flash attention backward on tcgen05 (head_dim 64, bf16 operands, fp32 accumulation in TMEM).
// Gradient of F.scaled_dot_product_attention (osu_dreamer/common/attn.py:82) with recomputed probabilities:
//   P = exp(S*scale - lse),  dP = dO V^T,  dS = P o (dP - D) * scale,  D = rowsum(dO o O)
//   dV = P^T dO,  dK = dS^T Q,  dQ = dS K
// Two kernels so that no gradient needs cross-CTA atomics:
//   attn_bwd_dq_kernel   : CTA = one 128-row q tile, loops over 64-row kv tiles   (thread = q row)
//   attn_bwd_dkdv_kernel : CTA = one 128-row kv tile, loops over 64-row q tiles   (thread = kv row)
// All operand tiles come straight from the token-major qkv / dy buffers through 3-D tensor maps; the
// transposed uses (V, dO, Q, K as the B operand of the second GEMM) are MN-major UMMA descriptors on the
// same smem bytes, so nothing is ever transposed in memory.  Two CTAs per SM (256 TMEM columns each).
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

static constexpr int BW_THREADS = 192;
static constexpr int T128 = 128 * 128;  // bytes of a [128 x 64] bf16 tile
static constexpr int T64 = 64 * 128;    // bytes of a [64 x 64] bf16 tile
static constexpr uint32_t BW_TMEM_COLS = 256;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// write 32 consecutive bf16 of row `row` (columns c0..c0+31 of a 64-column K-major SW128 tile)
__device__ __forceinline__ void st_row32_sw128(uint32_t tile_base, int row, int c0, const float (&v)[32]) {
  const uint32_t rb = tile_base + row * 128;
  const int sw = row & 7;
#pragma unroll
  for (int u4 = 0; u4 < 4; ++u4) {
    const int u = (c0 >> 3) + u4;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rb + ((u ^ sw) << 4)),
                 "r"(pack_bf16(v[u4 * 8 + 0], v[u4 * 8 + 1])), "r"(pack_bf16(v[u4 * 8 + 2], v[u4 * 8 + 3])),
                 "r"(pack_bf16(v[u4 * 8 + 4], v[u4 * 8 + 5])), "r"(pack_bf16(v[u4 * 8 + 6], v[u4 * 8 + 7]))
                 : "memory");
  }
}

struct AttnBwdParams {
  CUtensorMap tma_qkv128;  // qkv  dims (3*dh, L, B), box (64, 128, 1)
  CUtensorMap tma_qkv64;   //                         box (64, 64, 1)
  CUtensorMap tma_dy128;   // dy   dims (dh, L, B),   box (64, 128, 1)
  CUtensorMap tma_dy64;    //                         box (64, 64, 1)
  const float* lse;        // [B, H, L]
  const float* dsum;       // [B, H, L]  D = rowsum(dO o O)
  __nv_bfloat16* dqkv;     // [B*L, 3*dh]  (dq | dk | dv), gradients w.r.t. the roped q, k and v
  int B, H, L, dh;
  float scale, scale_log2;
  const int* gate;  // nullable: the kernels return immediately when *gate == 0 (fallback launches of the fused path)
};

// ================================================================================================= dQ
//   smem : Q [128x64] | dO [128x64] (staging only) | K_j x2 [64x64] | V_j x2 [64x64] | barriers
//   TMEM : S cols [0,64) | dP [64,128) | dQ [128,192) | Q bf16 [192,224) | dO bf16 [224,256)
//          Q and dO are copied once into TMEM and used from there as the A operands of S = Q K_j^T and
//          dP = dO V_j^T (a 128x64 smem A tile would otherwise be re-read for every kv tile and make those MMAs
//          shared-memory-bound); dS (bf16) is written over dP's first 32 columns and consumed from there by
//          dQ += dS K_j, which is therefore issued before the next S / dP MMAs (in-order tensor pipe).
static constexpr int DQ_SMEM_TILES = 2 * T128 + 4 * T64;
static constexpr int DQ_SMEM_BYTES = DQ_SMEM_TILES + 256;

__global__ void __launch_bounds__(BW_THREADS, 2) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (p.gate != nullptr && *p.gate == 0) return;
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + DQ_SMEM_TILES + 128 > dyn) __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + T128;
  uint8_t* sK = sDO + T128;     // 2 stages
  uint8_t* sV = sK + 2 * T64;   // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * T64);
  uint64_t* qdo_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* dq_done = bars + 7;
  uint64_t* qt_ready = bars + 8;
  uint64_t* dp_full = bars + 9;   // dP_j complete
  uint64_t* p1_done = bars + 10;  // softmax has consumed S_j (phase 1 finished)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (p.L + 127) / 128;
  // Work items = (q tile, head, sample).  Launched directly the grid has one CTA per item and this loop runs once; as the
  // gated fallback of the single-pass backward the grid is two CTAs per SM (an empty 16 384-CTA launch cost 70 us, twice per
  // layer), and in the rare case that the gate is open every CTA walks its items, re-initialising barriers and TMEM each time.
  const unsigned n_work = (unsigned)n_qt * p.H * p.B;
  if (warp == 1) {  // TMEM: once per CTA (the allocation permit is relinquished, so not per item)
    tmem_alloc(tmem_slot, BW_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#pragma unroll 1
  for (unsigned work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int qt = work % n_qt;
  const int bh = work / n_qt;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * 128;
  const int n_kv = (p.L + 63) / 64;

  if (warp == 0 && lane == 0) {
    mbar_init(qdo_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(ds_full, 4);
    mbar_init(dq_done, 1);
    mbar_init(qt_ready, 4);
    mbar_init(dp_full, 1);
    mbar_init(p1_done, 4);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(qdo_full, 2 * T128);
      tma_load_3d(sQ, &p.tma_qkv128, qdo_full, h * 64, q0, b);
      tma_load_3d(sDO, &p.tma_dy128, qdo_full, h * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * T64);
        tma_load_3d(sK + st * T64, &p.tma_qkv64, &kv_full[st], p.dh + h * 64, j * 64, b);
        tma_load_3d(sV + st * T64, &p.tma_qkv64, &kv_full[st], 2 * p.dh + h * 64, j * 64, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id_s = make_idesc(FMT_BF16, 0, 0, 128, 64);   // S, dP : K-major x K-major
      const uint32_t id_o = make_idesc(FMT_BF16, 0, 1, 128, 64);   // dQ    : dS K-major x K_j MN-major
      const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;
      const uint32_t tQ = tmem_base + 192, tDO = tmem_base + 224, tDS = tDP;
      // Software pipeline (per kv tile j):  S_j | dP_j | S_{j+1} (once phase 1 of j is done) | dQ_j (needs dS_j) | dP_{j+1}
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * T64);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tS, tQ + k * 8, make_smem_desc(aK + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(s_full);
      };
      auto issue_dp = [&](int j) {
        const uint32_t aV = smem_u32(sV + (j & 1) * T64);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tDP, tDO + k * 8, make_smem_desc(aV + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(dp_full);
      };
      mbar_wait(qt_ready, 0);  // Q / dO copied into TMEM by the softmax warps
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        if (j + 1 < n_kv) {
          mbar_wait(p1_done, j & 1);  // S_j consumed
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(ds_full, j & 1);  // dS_j written over dP_j
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * T64);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tDQ, tDS + k * 8, make_smem_desc(aK + k * 16 * 128, 0, 1024), id_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(dq_done);
        if (j + 1 < n_kv) issue_dp(j + 1);  // overwrites dS_j only after dQ_j has read it
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off, tDP = tmem_base + 64 + lane_off, tDQ = tmem_base + 128 + lane_off;
    const uint32_t tDS = tDP;
    {  // one-time copy of this thread's Q and dO rows (128 B each, SW128 smem) into TMEM columns [192,256)
      mbar_wait(qdo_full, 0);
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        const uint32_t base = smem_u32(which == 0 ? sQ : sDO) + row * 128;
        uint32_t r[32];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(r[4 * u]), "=r"(r[4 * u + 1]), "=r"(r[4 * u + 2]), "=r"(r[4 * u + 3])
                       : "r"(base + ((u ^ (row & 7)) << 4)));
        __syncwarp();
        tmem_st32(tmem_base + 192 + which * 32 + lane_off, r);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qt_ready);
    }
    const int q = q0 + row;
    const bool ok = q < p.L;
    const size_t sidx = ((size_t)b * p.H + h) * p.L + (ok ? q : 0);
    const float lse2 = ok ? p.lse[sidx] * 1.4426950408889634f : INFINITY;
    const float dsum = ok ? p.dsum[sidx] : 0.f;
    const float c = p.scale_log2;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.L - j * 64;
      // ---- phase 1: P = exp2(S * c - lse2)  (while the dP MMA is in flight)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float pr[64];
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t rs[32];
        __syncwarp();
        tmem_ld32(tS + cch * 32, rs);
        tmem_wait_ld();
        const float2 c2 = make_float2(c, c), nl2 = make_float2(-lse2, -lse2);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 a = ffma2(make_float2(__uint_as_float(rs[i]), __uint_as_float(rs[i + 1])), c2, nl2);
          pr[cch * 32 + i] = ex2f(a.x);
          pr[cch * 32 + i + 1] = ex2f(a.y);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p1_done);
      // ---- phase 2: dS = P o (dP - D) * scale -> bf16 over dP's first 32 columns
      mbar_wait(dp_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t rp[32];
        __syncwarp();
        tmem_ld32(tDP + cch * 32, rp);
        tmem_wait_ld();
        uint32_t pk[16];
        const float2 nd2 = make_float2(-dsum, -dsum);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // dS / scale = P o (dP - D); the 1/sqrt(d) factor is applied once to the dQ accumulator in the epilogue
          float2 d = fmul2(make_float2(pr[cch * 32 + 2 * i], pr[cch * 32 + 2 * i + 1]),
                           fadd2(make_float2(__uint_as_float(rp[2 * i]), __uint_as_float(rp[2 * i + 1])), nd2));
          if (valid < 64) {
            if (cch * 32 + 2 * i >= valid) d.x = 0.f;
            if (cch * 32 + 2 * i + 1 >= valid) d.y = 0.f;
          }
          pk[i] = pack_bf16(d.x, d.y);
        }
        tmem_st16(tDS + cch * 16, pk);  // over dP columns [16*cch, 16*cch+16): already consumed
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_done, (n_kv - 1) & 1);
    tc_fence_after();
#pragma unroll 1
    for (int cch = 0; cch < 2; ++cch) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tDQ + cch * 32, r);
      tmem_wait_ld();
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + ((size_t)b * p.L + q) * (3 * p.dh) + h * 64 + cch * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_uint4(pack_bf16(__uint_as_float(r[8 * i]) * p.scale, __uint_as_float(r[8 * i + 1]) * p.scale),
                              pack_bf16(__uint_as_float(r[8 * i + 2]) * p.scale, __uint_as_float(r[8 * i + 3]) * p.scale),
                              pack_bf16(__uint_as_float(r[8 * i + 4]) * p.scale, __uint_as_float(r[8 * i + 5]) * p.scale),
                              pack_bf16(__uint_as_float(r[8 * i + 6]) * p.scale, __uint_as_float(r[8 * i + 7]) * p.scale));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (work + gridDim.x < n_work) {  // another item follows: every thread is past its last wait (barrier above)
    if (warp == 0 && lane == 0)
      for (int i = 0; i < 11; ++i) mbar_inval(&bars[i]);
    __syncthreads();
  }
  }  // work loop
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, BW_TMEM_COLS);
  }
}

// ================================================================================================= dK, dV
//   smem : K [128x64] | V [128x64] | Q_i x2 [64x64] | dO_i x2 [64x64] | lse2/D staging [2][2][64] f32 | barriers
//   TMEM : S^T cols [0,64) | dP^T [64,128) | dV [128,192) | dK [192,256); the bf16 P^T / dS^T tiles are written
//          back over the first 32 columns of S^T / dP^T and consumed from there as the A operands of
//          dV += P^T dO_i and dK += dS^T Q_i (no shared-memory staging; the tensor pipe is in order, so the next
//          S^T / dP^T MMAs are simply issued after them).
static constexpr int DKV_SMEM_TILES = 2 * T128 + 4 * T64;
static constexpr int DKV_SMEM_BYTES = DKV_SMEM_TILES + 1024 + 256;

__global__ void __launch_bounds__(BW_THREADS, 2) attn_bwd_dkdv_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (p.gate != nullptr && *p.gate == 0) return;
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + DKV_SMEM_TILES + 1024 + 128 > dyn) __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = sK + T128;
  uint8_t* sQ = sV + T128;       // 2 stages [64 x 64]
  uint8_t* sDO = sQ + 2 * T64;   // 2 stages
  float* sStat = reinterpret_cast<float*>(sDO + 2 * T64);  // [2 buf][lse2 64 | D 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDO + 2 * T64 + 1024);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // S^T_i complete
  uint64_t* ds_full = bars + 6;   // dS^T_i written (TMEM)
  uint64_t* acc_done = bars + 7;  // dK_i issued-and-complete (everything before it too)
  uint64_t* dp_full = bars + 8;   // dP^T_i complete
  uint64_t* pt_full = bars + 9;   // P^T_i written (TMEM)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kt = (p.L + 127) / 128;
  const unsigned n_work = (unsigned)n_kt * p.H * p.B;  // see attn_bwd_dq_kernel
  if (warp == 1) {
    tmem_alloc(tmem_slot, BW_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#pragma unroll 1
  for (unsigned work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int kt = work % n_kt;
  const int bh = work / n_kt;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kt * 128;
  const int n_q = (p.L + 63) / 64;

  if (warp == 0 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(ds_full, 4);
    mbar_init(acc_done, 1);
    mbar_init(dp_full, 1);
    mbar_init(pt_full, 4);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * T128);
      tma_load_3d(sK, &p.tma_qkv128, kv_full, p.dh + h * 64, kv0, b);
      tma_load_3d(sV, &p.tma_qkv128, kv_full, 2 * p.dh + h * 64, kv0, b);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * T64);
        tma_load_3d(sQ + st * T64, &p.tma_qkv64, &q_full[st], h * 64, i * 64, b);
        tma_load_3d(sDO + st * T64, &p.tma_dy64, &q_full[st], h * 64, i * 64, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id_s = make_idesc(FMT_BF16, 0, 0, 128, 64);  // S^T = K Q_i^T, dP^T = V dO_i^T
      const uint32_t id_o = make_idesc(FMT_BF16, 0, 1, 128, 64);  // dV = P^T dO_i, dK = dS^T Q_i (B MN-major)
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 192;
      // Software pipeline (per q tile i):   S^T_i | dP^T_i | dV_i (needs P^T_i) | S^T_{i+1} | dK_i (needs dS^T_i) | dP^T_{i+1}
      // so that the exp phase of the softmax overlaps the dP^T MMA, the dS phase overlaps dV and the next S^T,
      // and the softmax warps never idle in steady state.  The bf16 P^T / dS^T tiles alias the first 32 columns of
      // S^T / dP^T, which the in-order tensor pipe makes safe with exactly this issue order.
      auto issue_st = [&](int i) {
        const int st = i & 1;
        mbar_wait(&q_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + st * T64);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tS, make_smem_desc(aK + k * 32, 0, 1024), make_smem_desc(aQ + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(s_full);
      };
      auto issue_dpt = [&](int i) {
        const uint32_t aDO = smem_u32(sDO + (i & 1) * T64);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tDP, make_smem_desc(aV + k * 32, 0, 1024), make_smem_desc(aDO + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(dp_full);
      };
      mbar_wait(kv_full, 0);
      issue_st(0);
      issue_dpt(0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        const uint32_t aQ = smem_u32(sQ + st * T64), aDO = smem_u32(sDO + st * T64);
        mbar_wait(pt_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tDV, tS + k * 8, make_smem_desc(aDO + k * 16 * 128, 0, 1024), id_o, (i > 0 || k > 0) ? 1u : 0u);
        if (i + 1 < n_q) issue_st(i + 1);  // overwrites P^T_i only after dV_i has read it
        mbar_wait(ds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tDK, tDP + k * 8, make_smem_desc(aQ + k * 16 * 128, 0, 1024), id_o, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&q_empty[st]);
        umma_commit(acc_done);
        if (i + 1 < n_q) issue_dpt(i + 1);  // overwrites dS^T_i only after dK_i has read it
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // kv row
    const int tid = threadIdx.x - 64;  // 0..127 within the softmax group
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off, tDP = tmem_base + 64 + lane_off;
    const uint32_t tDV = tmem_base + 128 + lane_off, tDK = tmem_base + 192 + lane_off;
    const float c = p.scale_log2;
    const size_t sbase = ((size_t)b * p.H + h) * p.L;
    // lse2 / D of a q tile are staged in smem (double-buffered); the global loads for tile i+1 are issued before
    // tile i is processed so that their latency is off the critical path.  q rows past L: lse2 = +inf -> P = 0
    auto load_stat = [&](int i) -> float {
      const int qi = i * 64 + (tid & 63);
      const bool okq = qi < p.L;
      if (tid < 64) return okq ? -p.lse[sbase + qi] * 1.4426950408889634f : -INFINITY;  // stored negated
      return okq ? -p.dsum[sbase + qi] : 0.f;  // stored negated: dS uses dP + (-D)
    };
    sStat[tid] = load_stat(0);
    named_bar_sync(1, 128);
    for (int i = 0; i < n_q; ++i) {
      float* st = sStat + (i & 1) * 128;
      float next_stat = 0.f;
      if (i + 1 < n_q) next_stat = load_stat(i + 1);
      // ---- phase 1: P^T = exp2(S^T * c - lse2[q])  (needs only S^T; runs while the dP^T MMA is in flight)
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      float pt[64];
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t rs[32];
        __syncwarp();
        tmem_ld32(tS + cch * 32, rs);
        tmem_wait_ld();
        const float4* l4 = reinterpret_cast<const float4*>(st + cch * 32);  // broadcast LDS.128
        const float2 c2 = make_float2(c, c);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 lv = l4[k4];  // -lse2 of 4 consecutive q columns (stored negated)
          const float2 a = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 0]), __uint_as_float(rs[k4 * 4 + 1])), c2,
                                 make_float2(lv.x, lv.y));
          const float2 b = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 2]), __uint_as_float(rs[k4 * 4 + 3])), c2,
                                 make_float2(lv.z, lv.w));
          pt[cch * 32 + k4 * 4 + 0] = ex2f(a.x);
          pt[cch * 32 + k4 * 4 + 1] = ex2f(a.y);
          pt[cch * 32 + k4 * 4 + 2] = ex2f(b.x);
          pt[cch * 32 + k4 * 4 + 3] = ex2f(b.y);
        }
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) pk[k] = pack_bf16(pt[cch * 32 + 2 * k], pt[cch * 32 + 2 * k + 1]);
        tmem_st16(tS + cch * 16, pk);  // over S^T columns already consumed
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pt_full);
      // ---- phase 2: dS^T = P^T o (dP^T - D[q]) * scale  (runs while dV_i and S^T_{i+1} are in flight)
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t rp[32];
        __syncwarp();
        tmem_ld32(tDP + cch * 32, rp);
        tmem_wait_ld();
        const float4* d4 = reinterpret_cast<const float4*>(st + 64 + cch * 32);
        uint32_t pk[16];
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 dv = d4[k4];  // -D of 4 consecutive q columns
          // dS^T / scale = P^T o (dP^T - D); 1/sqrt(d) is applied to the dK accumulator in the epilogue
          const float2 e0 = fmul2(make_float2(pt[cch * 32 + k4 * 4 + 0], pt[cch * 32 + k4 * 4 + 1]),
                                  fadd2(make_float2(__uint_as_float(rp[k4 * 4 + 0]), __uint_as_float(rp[k4 * 4 + 1])),
                                        make_float2(dv.x, dv.y)));
          const float2 e1 = fmul2(make_float2(pt[cch * 32 + k4 * 4 + 2], pt[cch * 32 + k4 * 4 + 3]),
                                  fadd2(make_float2(__uint_as_float(rp[k4 * 4 + 2]), __uint_as_float(rp[k4 * 4 + 3])),
                                        make_float2(dv.z, dv.w)));
          pk[k4 * 2] = pack_bf16(e0.x, e0.y);
          pk[k4 * 2 + 1] = pack_bf16(e1.x, e1.y);
        }
        tmem_st16(tDP + cch * 16, pk);  // over dP^T columns already consumed
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (i + 1 < n_q) {
        sStat[((i + 1) & 1) * 128 + tid] = next_stat;  // buffer (i+1)&1 was last read during tile i-1
        named_bar_sync(1, 128);
      }
    }
    mbar_wait(acc_done, (n_q - 1) & 1);
    tc_fence_after();
    const int kv = kv0 + row;
    const bool ok = kv < p.L;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {  // 0: dK -> column block 1, 1: dV -> column block 2
#pragma unroll 1
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32((which == 0 ? tDK : tDV) + cch * 32, r);
        tmem_wait_ld();
        if (which == 0) {
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * p.scale);
        }
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.dqkv + ((size_t)b * p.L + kv) * (3 * p.dh) + (1 + which) * p.dh +
                                                h * 64 + cch * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            dst[k] = make_uint4(pack_bf16(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])),
                                pack_bf16(__uint_as_float(r[8 * k + 2]), __uint_as_float(r[8 * k + 3])),
                                pack_bf16(__uint_as_float(r[8 * k + 4]), __uint_as_float(r[8 * k + 5])),
                                pack_bf16(__uint_as_float(r[8 * k + 6]), __uint_as_float(r[8 * k + 7])));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (work + gridDim.x < n_work) {  // another item follows: every thread is past its last wait (barrier above)
    if (warp == 0 && lane == 0)
      for (int i = 0; i < 10; ++i) mbar_inval(&bars[i]);
    __syncthreads();
  }
  }  // work loop
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, BW_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------- D = rowsum(dO o O)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                                     float* __restrict__ dsum, int B, int H, int L) {
  // warp per token: lane covers 32 contiguous columns of the 1024 (= half a head)
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= B * L) return;
  const uint4* a = reinterpret_cast<const uint4*>(y + (size_t)t * (H * 64) + lane * 32);
  const uint4* g = reinterpret_cast<const uint4*>(dy + (size_t)t * (H * 64) + lane * 32);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 av = a[i], gv = g[i];
    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&av);
    const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc = fmaf(__low2float(ap[k]), __low2float(gp[k]), acc);
      acc = fmaf(__high2float(ap[k]), __high2float(gp[k]), acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if ((lane & 1) == 0) {
    const int hh = lane >> 1;
    const int b = t / L, l = t % L;
    dsum[((size_t)b * H + hh) * L + l] = acc;
  }
}

int launch_attn_bwd_prep(const void* y, const void* dy, float* dsum, int B, int L, int H, cudaStream_t stream) {
  attn_bwd_prep_kernel<<<ceil_div(B * L, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y),
                                                               static_cast<const __nv_bfloat16*>(dy), dsum, B, H, L);
  OSD_LAUNCHED();
  return 0;
}

static int launch_attn_bwd_main(const void* qkv, const void* dy, const float* lse, const float* dsum, void* dqkv,
                                int B, int L, int H, const int* gate, cudaStream_t stream);

int launch_attn_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* dsum, void* dqkv, int B,
                    int L, int H, cudaStream_t stream) {
  OSD_CHECK(qkv && y && dy && lse && dsum && dqkv && B > 0 && L > 0 && H == 16, "attn_bwd: bad arguments");
  attn_bwd_prep_kernel<<<ceil_div(B * L, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y),
                                                               static_cast<const __nv_bfloat16*>(dy), dsum, B, H, L);
  OSD_LAUNCHED();
  return launch_attn_bwd_main(qkv, dy, lse, dsum, dqkv, B, L, H, nullptr, stream);
}

// the two kernels with dsum = rowsum(dO o O) already computed; they do nothing when *gate == 0
int launch_attn_bwd_gated(const void* qkv, const void* dy, const float* lse, const float* dsum, void* dqkv, int B, int L,
                          int H, const int* gate, cudaStream_t stream) {
  return launch_attn_bwd_main(qkv, dy, lse, dsum, dqkv, B, L, H, gate, stream);
}

static int launch_attn_bwd_main(const void* qkv, const void* dy, const float* lse, const float* dsum, void* dqkv,
                                int B, int L, int H, const int* gate, cudaStream_t stream) {
  const int dh = H * 64;
  AttnBwdParams p;
  p.gate = gate;
  {
    uint64_t dims[3] = {(uint64_t)3 * dh, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)3 * dh * 2, (uint64_t)L * 3 * dh * 2};
    uint32_t box128[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    OSD_TRY(make_tmap(&p.tma_qkv128, qkv, 2, 3, dims, strides, box128));
    OSD_TRY(make_tmap(&p.tma_qkv64, qkv, 2, 3, dims, strides, box64));
  }
  {
    uint64_t dims[3] = {(uint64_t)dh, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)dh * 2, (uint64_t)L * dh * 2};
    uint32_t box128[3] = {64, 128, 1}, box64[3] = {64, 64, 1};
    OSD_TRY(make_tmap(&p.tma_dy128, dy, 2, 3, dims, strides, box128));
    OSD_TRY(make_tmap(&p.tma_dy64, dy, 2, 3, dims, strides, box64));
  }
  p.lse = lse;
  p.dsum = dsum;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.B = B; p.H = H; p.L = L; p.dh = dh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM_BYTES));
    OSD_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM_BYTES));
  }
  long long grid = (long long)ceil_div(L, 128) * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_bwd: grid too large");
  if (gate != nullptr && grid > 2 * num_sms()) grid = 2 * num_sms();  // fallback launch: resident CTAs only, each walks its items
  attn_bwd_dkdv_kernel<<<(unsigned)grid, BW_THREADS, DKV_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  attn_bwd_dq_kernel<<<(unsigned)grid, BW_THREADS, DQ_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
