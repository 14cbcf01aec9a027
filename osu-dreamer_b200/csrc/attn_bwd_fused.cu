// Single-pass flash attention backward on tcgen05 (head_dim 64, bf16 operands, fp32 accumulation in TMEM).
// Gradient of F.scaled_dot_product_attention (osu_dreamer/common/attn.py:82):
//   P = exp(S*scale - lse),  dP = dO V^T,  dS = P o (dP - D) * scale,  D = rowsum(dO o O)
//   dV = P^T dO,  dK = dS^T Q,  dQ = dS K
// One kernel, five GEMMs and ONE exponential per score element (the two-kernel path in attn_bwd.cu recomputes S
// and P in both kernels: seven GEMMs, two exponentials).  CTA = one 128-row kv tile of one (batch, head); it
// loops over all 128-row q tiles, keeps dK / dV in TMEM, and adds its [128 x 64] fp32 partial of dQ into a
// global fp32 accumulator with TMA reduce-add (cp.reduce.async.bulk.tensor ... .add), as many flash-attention
// backward kernels do.  Each CTA starts its q loop at a different tile so that the 64 CTAs of one (b, h) do not
// hit the same dQ rows at the same time.
//
//   warps    : 0 TMA producer | 1 TMEM allocator + UMMA issuer (event-driven) | 2-9 softmax (2 column groups x 4
//              lane quadrants, thread = kv row, separate barriers per group) | 10-13 dQ drain (TMEM -> swizzled smem
//              -> TMA reduce-add; warp = 32 q rows x 64 d columns)
//   TMEM     : S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ_i [384,448) | K bf16 [448,480) |
//              V bf16 [480,512).  K and V are the A operands of S^T = K Q_i^T and dP^T = V dO_i^T straight from TMEM;
//              the bf16 P^T / dS^T tiles are written back over the consumed S^T / dP^T columns (column group g at
//              +64g) and are the TMEM A operands of dV += P^T dO_i and dK += dS^T Q_i.
//   smem     : K | V (staging for the TMEM copy, then dQ staging; K also B operand of dQ) | Q_i x3 | dO_i x3 |
//              dS^T x2 (bf16, MN-major A operand of dQ_i = dS_i K) | -lse2 / -D x3 (bulk-copied with Q_i) | barriers
#include "kernels.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace osd {

static constexpr int FB_THREADS = 320;
#ifndef OSD_FB_EMU
#define OSD_FB_EMU 2
#endif
static constexpr int FB_EMU = OSD_FB_EMU;  // exponentials computed on the FMA pipe: 0 none, 1 a quarter, 2 half
static constexpr int FT = 128 * 128;  // bytes of a [128 x 64] bf16 tile
static constexpr uint32_t FB_TMEM_COLS = 512;
static constexpr int FB_STAT_BYTES = 2 * 256 * 4;
static constexpr int FB_STG_BYTES = 8 * 4096;  // 8 softmax warps x [32 rows x 128 B]
static constexpr int FB_SMEM_TILES = 8 * FT + FB_STG_BYTES + FB_STAT_BYTES;
static constexpr int FB_SMEM_BYTES = FB_SMEM_TILES + 256 + 1024;

__device__ __forceinline__ float fb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnBwdFusedParams {
  CUtensorMap tma_qkv;  // qkv    dims (3*dh, L, B), box (64, 128, 1)
  CUtensorMap tma_dy;   // w-scaled dO: dims (dh, L, B), box (64, 128, 1)
  CUtensorMap tma_dq;   // dq_acc dims (dh, L, B) fp32, box (32, 32, 1)
  const float* mtile;   // [B*H, n_t]  m = max over the q tile of lse * log2(e)
  const float* dneg;    // [B*H, Lp]   -w D,  w = 2^(m - lse2), D = rowsum(dO o O); zero past L (Lp = L rounded up to 128)
  const int* fallback;  // != 0: the statistics of some q tile span too many octaves for the w-scaling -> do nothing
  __nv_bfloat16* dqkv;  // [B*L, 3*dh]: this kernel writes the dk and dv column blocks
  int B, H, L, Lp, dh;
  float scale, scale_log2;
  int debug_skip;  // timing experiments only (OSD_FB_SKIP): 1 = no dQ reduce-add (results wrong)
  unsigned long long* trace;  // nullable (tools/trace_attn_bwd.py): event timeline of one CTA, 3 writers x 1024 records
  int trace_cta;
};

// timeline record: (event << 48) | (tile << 32) | low 32 bits of the SM clock
#define FB_TRACE(slot, ev, tile)                                                                                     \
  do {                                                                                                               \
    if (tr != nullptr && tr_n < 1024)                                                                                \
      tr[(slot) * 1024 + tr_n++] = ((unsigned long long)(ev) << 48) | ((unsigned long long)(tile) << 32) |           \
                                   (unsigned long long)(clock64() & 0xffffffffll);                                   \
  } while (0)

// W16 (OSD_FB_W16=1, A/B): SIXTEEN softmax warps -- four q-column groups of 32 instead of two of 64, four warps per
// sub-partition instead of two.  The forward showed (tools/micro/softmax_bench.cu, attn_fwd_pp3.cu) that the per-element
// softmax chain is latency bound at two warps per sub-partition; this kernel's softmax warps are its critical path (§5: they
// never idle, the tensor pipe is 47 % busy).  Differences: a thread owns 32 score columns; P~ is carried from phase 1 to phase 2
// as the bf16 pairs that were written to TMEM (16 registers instead of 64 fp32: the register budget at 18 warps is 96), i.e.
// dS uses the same rounded P~ as dV; P~^T / dS^T of group g sit in the first 16 columns of the group's own 32 consumed columns;
// the dQ drain stays with groups 0 / 1 (8 warps x [32 x 32] fp32 as before).
template <bool W16>
__global__ void __launch_bounds__(W16 ? 576 : FB_THREADS, 1) attn_bwd_fused_kernel(const __grid_constant__ AttnBwdFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (*p.fallback != 0) return;
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + FB_SMEM_TILES + 256 > dyn) __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = sK + FT;
  uint8_t* sQ = sV + FT;        // 2 stages [128 x 64]
  uint8_t* sDO = sQ + 2 * FT;   // 2 stages
  uint8_t* sDS = sDO + 2 * FT;  // [2 q-chunks][128 kv rows][64 q] bf16
  uint8_t* sStg = sDS + 2 * FT;
  float* sStat = reinterpret_cast<float*>(sStg + FB_STG_BYTES);  // [2 buf][-lse2 128 | -D 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + FB_STG_BYTES + FB_STAT_BYTES);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // [2]
  uint64_t* q_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;    // S^T_i complete
  uint64_t* dp_full = bars + 6;   // dP^T_i complete
  uint64_t* pt_full = bars + 7;   // P^T_i written (TMEM)
  uint64_t* ds_full = bars + 8;   // dS^T_i written (TMEM + smem)
  uint64_t* dq_full = bars + 9;   // dQ_i complete (also: dS^T smem tile consumed)
  uint64_t* dq_empty = bars + 10; // dQ_i drained out of TMEM
  uint64_t* kvt_ready = bars + 11;  // K / V copied into TMEM
  uint64_t* acc_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  uint64_t* st_empty = bars + 14;  // [2] -w D statistics of the stage read by all softmax warps (generic-proxy reads vs the
                                   // next bulk copy into the same stage; also implied by ds_full -> dK -> q_empty, but that
                                   // chain runs through tcgen05.commit, which compute-sanitizer racecheck cannot follow)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_t = (p.L + 127) / 128;  // kv tiles == q tiles
  const int kt = blockIdx.x % n_t;
  const int bh = blockIdx.x / n_t;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kt * 128;
  const int n_q = n_t;
  const int i0 = kt;  // q-loop rotation: iteration i works on q tile (i0 + i) % n_q
  unsigned long long* tr = (p.trace != nullptr && (int)blockIdx.x == p.trace_cta && (threadIdx.x & 31) == 0) ? p.trace : nullptr;
  int tr_n = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    tma_prefetch_desc(&p.tma_dy);
    tma_prefetch_desc(&p.tma_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(pt_full, W16 ? 16 : 8);
    mbar_init(ds_full, W16 ? 16 : 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    mbar_init(kvt_ready, 8);
    mbar_init(acc_done, 1);
    mbar_init(&st_empty[0], W16 ? 16 : 8);
    mbar_init(&st_empty[1], W16 ? 16 : 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, FB_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * FT);
      tma_load_3d(sK, &p.tma_qkv, kv_full, p.dh + h * 64, kv0, b);
      tma_load_3d(sV, &p.tma_qkv, kv_full, 2 * p.dh + h * 64, kv0, b);
      int qi = i0;
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_wait(&st_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * FT + 512);
        tma_load_3d(sQ + st * FT, &p.tma_qkv, &q_full[st], h * 64, qi * 128, b);
        tma_load_3d(sDO + st * FT, &p.tma_dy, &q_full[st], h * 64, qi * 128, b);
        const size_t so = ((size_t)b * p.H + h) * p.Lp + (size_t)qi * 128;
        bulk_load_1d(sStat + st * 256 + 128, p.dneg + so, 512, &q_full[st]);
        if (++qi == n_q) qi = 0;
      }
    }
  } else if (warp == 1) {
    // ================================================================== UMMA issuer
    if (elect_one()) {
      const uint32_t id_s = make_idesc(FMT_BF16, 0, 0, 128, 128);  // S^T, dP^T: A in TMEM, B K-major, N = 128 q
      const uint32_t id_o = make_idesc(FMT_BF16, 0, 1, 128, 64);   // dV, dK  : A in TMEM, B MN-major, N = 64 d
      const uint32_t id_q = make_idesc(FMT_BF16, 1, 1, 128, 64);   // dQ      : A = dS MN-major smem, B = K MN-major
      const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320;
      const uint32_t tDQ = tmem_base + 384, tK = tmem_base + 448, tV = tmem_base + 480;
      const uint32_t aK = smem_u32(sK), aDS = smem_u32(sDS);
      auto issue_s = [&](int i) {
        const int st = i & 1;
        mbar_wait(&q_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + st * FT);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tS, tK + k * 8, make_smem_desc(aQ + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(s_full);
      };
      auto issue_dp = [&](int i) {
        const uint32_t aDO = smem_u32(sDO + (i & 1) * FT);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tDP, tV + k * 8, make_smem_desc(aDO + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(dp_full);
      };
      mbar_wait(kvt_ready, 0);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        const uint32_t aQ = smem_u32(sQ + st * FT), aDO = smem_u32(sDO + st * FT);
        const uint32_t acc = i > 0 ? 1u : 0u;
        mbar_wait(pt_full, i & 1);
        tc_fence_after();
        FB_TRACE(0, 10, i);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV += P^T_i dO_i : q (K dim) chunk k lives at columns (k/4)*64 + (k%4)*8
          umma_f16_ts(tDV, tS + (W16 ? (k >> 1) * 32 + (k & 1) * 8 : (k >> 2) * 64 + (k & 3) * 8),
                      make_smem_desc(aDO + k * 16 * 128, 0, 1024), id_o, (k > 0) ? 1u : acc);
        if (i + 1 < n_q) issue_s(i + 1);  // overwrites P^T_i only after dV_i has read it
        mbar_wait(ds_full, i & 1);
        tc_fence_after();
        FB_TRACE(0, 12, i);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ts(tDK, tDP + (W16 ? (k >> 1) * 32 + (k & 1) * 8 : (k >> 2) * 64 + (k & 3) * 8),
                      make_smem_desc(aQ + k * 16 * 128, 0, 1024), id_o, (k > 0) ? 1u : acc);
        umma_commit(&q_empty[st]);
        if (i + 1 < n_q) issue_dp(i + 1);  // overwrites dS^T_i (TMEM) only after dK_i has read it
        if (i > 0) {
          mbar_wait(dq_empty, (i - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ_i = dS_i K : contraction over the 128 kv rows, 16 per instruction
          umma_f16_ss(tDQ, make_smem_desc(aDS + k * 16 * 128, FT, 1024), make_smem_desc(aK + k * 16 * 128, 0, 1024), id_q,
                      k > 0);
        umma_commit(dq_full);
        FB_TRACE(0, 16, i);
      }
      umma_commit(acc_done);
    }
  } else if constexpr (W16) {
    // ================================================================== softmax, 16 warps (thread = kv row, 32 q columns)
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;   // q-column group: columns [32 grp, 32 grp + 32)
    const int row = quad * 32 + lane;  // kv row
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool drainer = grp < 2;      // groups 0 / 1 also copy K / V into TMEM, drain dQ and write dK / dV
    if (drainer) {  // one-time copy of this thread's K (grp 0) or V (grp 1) row (128 B, SW128 smem) into TMEM
      mbar_wait(kv_full, 0);
      const uint32_t base = smem_u32(grp == 0 ? sK : sV) + row * 128;
      uint32_t r[32];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[4 * u]), "=r"(r[4 * u + 1]), "=r"(r[4 * u + 2]), "=r"(r[4 * u + 3])
                     : "r"(base + ((u ^ (row & 7)) << 4)));
      __syncwarp();
      tmem_st32(tmem_base + 448 + grp * 32 + lane_off, r);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(kvt_ready);
    }
    const uint32_t tS = tmem_base + lane_off + grp * 32, tDP = tmem_base + 128 + lane_off + grp * 32;
    const uint32_t ds_row = smem_u32(sDS) + (grp >> 1) * FT + row * 128;  // [q chunk of 64][kv row][64 q]: this group's 64-byte half
    const int cb = (grp & 1) * 4;
    const int sw = row & 7;
    const float c = p.scale_log2;
    const uint32_t tDQ = tmem_base + 384 + lane_off + grp * 32;
    uint8_t* stg = sStg + ((warp - 2) & 7) * 4096;
    const uint32_t stg_row = smem_u32(stg) + lane * 128;
    auto drain_load = [&](int j) {
      mbar_wait(dq_full, j & 1);
      tc_fence_after();
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tDQ, r);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dq_empty);
        tma_store_wait_read<0>();
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((u ^ (lane & 7)) << 4)), "r"(r[4 * u]),
                     "r"(r[4 * u + 1]), "r"(r[4 * u + 2]), "r"(r[4 * u + 3])
                     : "memory");
    };
    auto drain_issue = [&](int qj) {
      if (lane == 0 && !(p.debug_skip & 1)) {
        tma_reduce_add_3d(&p.tma_dq, stg, h * 64 + grp * 32, qj * 128 + quad * 32, b);
        tma_store_commit();
      }
    };
    int qt = i0, qprev = i0;
    const float* mrow = p.mtile + (size_t)bh * n_q;
    float m_cur = __ldg(mrow + qt);
    for (int i = 0; i < n_q; ++i) {
      const float* st = sStat + (i & 1) * 256 + grp * 32;
      int qn = qt + 1;
      if (qn == n_q) qn = 0;
      const float m_next = __ldg(mrow + qn);
      // ---- phase 1: P~^T = exp2(S^T * c - m)
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      uint32_t pk1[16];
      {
        uint32_t rs[32];
        __syncwarp();
        tmem_ld32(tS, rs);
        tmem_wait_ld();
        const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_cur, -m_cur);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float2 a = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 0]), __uint_as_float(rs[k4 * 4 + 1])), c2, nm2);
          const float2 bb = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 2]), __uint_as_float(rs[k4 * 4 + 3])), c2, nm2);
          const float2 e0 = make_float2(fb_ex2(a.x), fb_ex2(a.y));
          float2 e1;
          if (FB_EMU == 2 || (FB_EMU == 1 && (k4 & 1)))
            e1 = ex2_poly2(bb);
          else
            e1 = make_float2(fb_ex2(bb.x), fb_ex2(bb.y));
          pk1[2 * k4] = pack_bf16(e0.x, e0.y);
          pk1[2 * k4 + 1] = pack_bf16(e1.x, e1.y);
        }
      }
      __syncwarp();
      tmem_st16(tS, pk1);  // over the first 16 of this group's own 32 consumed columns
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pt_full);
      // ---- dQ_{i-1}: drained by groups 0 / 1; for everyone: it has consumed the dS^T smem tile
      if (i > 0) {
        if (drainer)
          drain_load(i - 1);
        else
          mbar_wait(dq_full, (i - 1) & 1);
      }
      // ---- phase 2: dS^T = P~^T o (dP~^T - D~[q])
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      uint32_t pk2[16];
      {
        uint32_t rp[32];
        __syncwarp();
        tmem_ld32(tDP, rp);
        tmem_wait_ld();
        const float4* d4 = reinterpret_cast<const float4*>(st + 128);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 dv = d4[k4];  // -D of 4 consecutive q columns
          const float2 p0 = make_float2(__uint_as_float(pk1[2 * k4] << 16), __uint_as_float(pk1[2 * k4] & 0xffff0000u));
          const float2 p1 = make_float2(__uint_as_float(pk1[2 * k4 + 1] << 16), __uint_as_float(pk1[2 * k4 + 1] & 0xffff0000u));
          const float2 e0 = fmul2(p0, fadd2(make_float2(__uint_as_float(rp[k4 * 4 + 0]), __uint_as_float(rp[k4 * 4 + 1])),
                                            make_float2(dv.x, dv.y)));
          const float2 e1 = fmul2(p1, fadd2(make_float2(__uint_as_float(rp[k4 * 4 + 2]), __uint_as_float(rp[k4 * 4 + 3])),
                                            make_float2(dv.z, dv.w)));
          pk2[2 * k4] = pack_bf16(e0.x, e0.y);
          pk2[2 * k4 + 1] = pack_bf16(e1.x, e1.y);
        }
      }
      __syncwarp();
      tmem_st16(tDP, pk2);  // TMEM copy: A operand of dK += dS^T Q_i
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4)  // smem copy: MN-major A operand of dQ_i = dS_i K
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + (((cb + u4) ^ sw) << 4)), "r"(pk2[4 * u4]),
                     "r"(pk2[4 * u4 + 1]), "r"(pk2[4 * u4 + 2]), "r"(pk2[4 * u4 + 3])
                     : "memory");
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(ds_full);
        mbar_arrive(&st_empty[i & 1]);
      }
      if (i > 0 && drainer) drain_issue(qprev);
      qprev = qt;
      qt = qn;
      m_cur = m_next;
    }
    if (drainer) {
      drain_load(n_q - 1);
      fence_proxy_async_smem();
      __syncwarp();
      drain_issue(qprev);
      // ---- epilogue: group 0 writes dK (x scale), group 1 writes dV
      mbar_wait(acc_done, 0);
      tc_fence_after();
      const int kv = kv0 + row;
      const bool ok = kv < p.L;
      const uint32_t tACC = tmem_base + (grp == 0 ? 320 : 256) + lane_off;
      const float mul = grp == 0 ? p.scale : 1.0f;
#pragma unroll 1
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tACC + cch * 32, r);
        tmem_wait_ld();
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.dqkv + ((size_t)b * p.L + kv) * (3 * p.dh) + (grp == 0 ? 1 : 2) * p.dh +
                                                h * 64 + cch * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            dst[k] = make_uint4(pack_bf16(__uint_as_float(r[8 * k]) * mul, __uint_as_float(r[8 * k + 1]) * mul),
                                pack_bf16(__uint_as_float(r[8 * k + 2]) * mul, __uint_as_float(r[8 * k + 3]) * mul),
                                pack_bf16(__uint_as_float(r[8 * k + 4]) * mul, __uint_as_float(r[8 * k + 5]) * mul),
                                pack_bf16(__uint_as_float(r[8 * k + 6]) * mul, __uint_as_float(r[8 * k + 7]) * mul));
        }
      }
      tc_fence_before();
      if (lane == 0) tma_store_wait<0>();
      __syncwarp();
    } else {
      tc_fence_before();
    }
  } else {
    // ================================================================== softmax (thread = kv row, 64 q columns)
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;   // q-column group: columns [64 grp, 64 grp + 64)
    const int row = quad * 32 + lane;  // kv row
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    {  // one-time copy of this thread's K (grp 0) or V (grp 1) row (128 B, SW128 smem) into TMEM
      mbar_wait(kv_full, 0);
      const uint32_t base = smem_u32(grp == 0 ? sK : sV) + row * 128;
      uint32_t r[32];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[4 * u]), "=r"(r[4 * u + 1]), "=r"(r[4 * u + 2]), "=r"(r[4 * u + 3])
                     : "r"(base + ((u ^ (row & 7)) << 4)));
      __syncwarp();
      tmem_st32(tmem_base + 448 + grp * 32 + lane_off, r);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(kvt_ready);
    }
    const uint32_t tS = tmem_base + lane_off + grp * 64, tDP = tmem_base + 128 + lane_off + grp * 64;
    const uint32_t ds_row = smem_u32(sDS) + grp * FT + row * 128;
    const int sw = row & 7;
    const float c = p.scale_log2;
    // dQ drain: this warp moves rows [32 quad, +32) x columns [32 grp, +32) of the finished dQ tile j (q tile qj)
    const uint32_t tDQ = tmem_base + 384 + lane_off + grp * 32;
    uint8_t* stg = sStg + (warp - 2) * 4096;
    const uint32_t stg_row = smem_u32(stg) + lane * 128;
    // part 1: TMEM -> registers -> staging smem (the proxy fence is shared with the dS^T smem writes of phase 2);
    // part 2 (after that fence): one lane issues the TMA reduce-add
    auto drain_load = [&](int j) {
      mbar_wait(dq_full, j & 1);
      tc_fence_after();
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tDQ, r);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dq_empty);
        tma_store_wait_read<0>();  // the previous reduce-add of this warp has read the staging tile
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + ((u ^ (lane & 7)) << 4)), "r"(r[4 * u]),
                     "r"(r[4 * u + 1]), "r"(r[4 * u + 2]), "r"(r[4 * u + 3])
                     : "memory");
    };
    auto drain_issue = [&](int qj) {
      if (lane == 0 && !(p.debug_skip & 1)) {
        tma_reduce_add_3d(&p.tma_dq, stg, h * 64 + grp * 32, qj * 128 + quad * 32, b);
        tma_store_commit();
      }
    };
    int qt = i0, qprev = i0;
    const float* mrow = p.mtile + (size_t)bh * n_q;
    float m_cur = __ldg(mrow + qt);
    const bool trw = quad == 2;  // traced warps: 2 (group 0) and 6 (group 1)
    for (int i = 0; i < n_q; ++i) {
      // -w D of this q tile, bulk-copied on the same barrier as Q_i / dO_i (complete before S^T_i was issued)
      const float* st = sStat + (i & 1) * 256 + grp * 64;
      int qn = qt + 1;
      if (qn == n_q) qn = 0;
      const float m_next = __ldg(mrow + qn);  // consumed one tile later
      // ---- phase 1: P~^T = exp2(S^T * c - m),  m = max lse2 of the q tile: no per-column statistic; the factor
      //      w[q] = 2^(m - lse2[q]) that turns P~ into P is folded into dO (dO~ = w dO) and D (D~ = w D)
      if (trw) FB_TRACE(1 + grp, 0, i);
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      if (trw) FB_TRACE(1 + grp, 1, i);
      float pt[64];
      uint32_t rsb[2][32];
      __syncwarp();
      tmem_ld32(tS, rsb[0]);
      tmem_wait_ld();
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        const uint32_t(&rs)[32] = rsb[cch];
        if (cch == 0) {
          __syncwarp();
          tmem_ld32(tS + 32, rsb[1]);  // second half streams out of TMEM while the first half is processed
        } else {
          tmem_wait_ld();
        }
        const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_cur, -m_cur);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float2 a = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 0]), __uint_as_float(rs[k4 * 4 + 1])), c2, nm2);
          const float2 bb = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 2]), __uint_as_float(rs[k4 * 4 + 3])), c2, nm2);
          pt[cch * 32 + k4 * 4 + 0] = fb_ex2(a.x);
          pt[cch * 32 + k4 * 4 + 1] = fb_ex2(a.y);
          if (FB_EMU == 2 || (FB_EMU == 1 && (k4 & 1))) {  // this pair on the FMA pipe instead of the SFU
            const float2 e = ex2_poly2(bb);
            pt[cch * 32 + k4 * 4 + 2] = e.x;
            pt[cch * 32 + k4 * 4 + 3] = e.y;
          } else {
            pt[cch * 32 + k4 * 4 + 2] = fb_ex2(bb.x);
            pt[cch * 32 + k4 * 4 + 3] = fb_ex2(bb.y);
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) pk[k] = pack_bf16(pt[cch * 32 + 2 * k], pt[cch * 32 + 2 * k + 1]);
        tmem_st16(tS + cch * 16, pk);  // over S^T columns this thread has already consumed
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pt_full);
      if (trw) FB_TRACE(1 + grp, 2, i);
      // ---- phase 2: dS^T = P~^T o (dP~^T - D~[q])   (1/sqrt(d) is applied to the dK / dQ results on the way out)
      if (i > 0) drain_load(i - 1);  // also: dQ_{i-1} has consumed the dS^T smem tile
      if (trw) FB_TRACE(1 + grp, 3, i);
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      if (trw) FB_TRACE(1 + grp, 4, i);
      uint32_t rpb[2][32];
      __syncwarp();
      tmem_ld32(tDP, rpb[0]);
      tmem_wait_ld();
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        const uint32_t(&rp)[32] = rpb[cch];
        if (cch == 0) {
          __syncwarp();
          tmem_ld32(tDP + 32, rpb[1]);
        } else {
          tmem_wait_ld();
        }
        const float4* d4 = reinterpret_cast<const float4*>(st + 128 + cch * 32);
        uint32_t pk[16];
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 dv = d4[k4];  // -D of 4 consecutive q columns
          const float2 e0 = fmul2(make_float2(pt[cch * 32 + k4 * 4 + 0], pt[cch * 32 + k4 * 4 + 1]),
                                  fadd2(make_float2(__uint_as_float(rp[k4 * 4 + 0]), __uint_as_float(rp[k4 * 4 + 1])),
                                        make_float2(dv.x, dv.y)));
          const float2 e1 = fmul2(make_float2(pt[cch * 32 + k4 * 4 + 2], pt[cch * 32 + k4 * 4 + 3]),
                                  fadd2(make_float2(__uint_as_float(rp[k4 * 4 + 2]), __uint_as_float(rp[k4 * 4 + 3])),
                                        make_float2(dv.z, dv.w)));
          pk[k4 * 2] = pack_bf16(e0.x, e0.y);
          pk[k4 * 2 + 1] = pack_bf16(e1.x, e1.y);
        }
        tmem_st16(tDP + cch * 16, pk);  // TMEM copy: A operand of dK += dS^T Q_i
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4)  // smem copy (row = kv, 64 q contiguous): MN-major A operand of dQ_i = dS_i K
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + (((cch * 4 + u4) ^ sw) << 4)),
                       "r"(pk[4 * u4]), "r"(pk[4 * u4 + 1]), "r"(pk[4 * u4 + 2]), "r"(pk[4 * u4 + 3])
                       : "memory");
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(ds_full);
        mbar_arrive(&st_empty[i & 1]);
      }
      if (trw) FB_TRACE(1 + grp, 5, i);
      if (i > 0) drain_issue(qprev);
      qprev = qt;
      qt = qn;
      m_cur = m_next;
    }
    drain_load(n_q - 1);
    fence_proxy_async_smem();
    __syncwarp();
    drain_issue(qprev);
    // ---- epilogue: group 0 writes dK (x scale), group 1 writes dV
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const int kv = kv0 + row;
    const bool ok = kv < p.L;
    const uint32_t tACC = tmem_base + (grp == 0 ? 320 : 256) + lane_off;
    const float mul = grp == 0 ? p.scale : 1.0f;
#pragma unroll 1
    for (int cch = 0; cch < 2; ++cch) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tACC + cch * 32, r);
      tmem_wait_ld();
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + ((size_t)b * p.L + kv) * (3 * p.dh) + (grp == 0 ? 1 : 2) * p.dh +
                                              h * 64 + cch * 32);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dst[k] = make_uint4(pack_bf16(__uint_as_float(r[8 * k]) * mul, __uint_as_float(r[8 * k + 1]) * mul),
                              pack_bf16(__uint_as_float(r[8 * k + 2]) * mul, __uint_as_float(r[8 * k + 3]) * mul),
                              pack_bf16(__uint_as_float(r[8 * k + 4]) * mul, __uint_as_float(r[8 * k + 5]) * mul),
                              pack_bf16(__uint_as_float(r[8 * k + 6]) * mul, __uint_as_float(r[8 * k + 7]) * mul));
      }
    }
    tc_fence_before();
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, FB_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------- statistics
// per (b, h, q tile): m = max lse2 over the tile; raises *fallback when the tile's lse2 values span more than 96
// octaves (then P~ = 2^(c s - m) could underflow for rows far below the maximum)
__global__ void attn_bwd_tilemax_kernel(const float* __restrict__ lse, float* __restrict__ mtile, int* __restrict__ fallback,
                                        int L, int n_t) {
  __shared__ float smx[4], smn[4];
  const int bh = blockIdx.x / n_t, t = blockIdx.x % n_t;
  const int l = t * 128 + threadIdx.x;
  const bool ok = l < L;
  const float v = ok ? lse[(size_t)bh * L + l] * 1.4426950408889634f : 0.f;
  float mx = ok ? v : -INFINITY, mn = ok ? v : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smx[threadIdx.x >> 5] = mx;
    smn[threadIdx.x >> 5] = mn;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = fmaxf(fmaxf(smx[0], smx[1]), fmaxf(smx[2], smx[3]));
    mn = fminf(fminf(smn[0], smn[1]), fminf(smn[2], smn[3]));
    mtile[blockIdx.x] = mx;
    if (!(mx - mn <= 96.f)) atomicOr(fallback, 1);  // also catches NaN / inf statistics
  }
}

// warp per padded token (lane = 32 contiguous columns of the 1024 = half a head):
//   D = rowsum(dO o O) -> dsum [B*H, L] (plain, for the two-pass fallback), dneg [B*H, Lp] = -w D (zero past L),
//   dys [B*L, 1024] = bf16(w dO),  w = 2^(m_tile - lse2)
__global__ void attn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                                      const float* __restrict__ lse, const float* __restrict__ mtile,
                                      float* __restrict__ dneg, float* __restrict__ dsum, __nv_bfloat16* __restrict__ dys,
                                      int B, int H, int L, int Lp) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= B * Lp) return;
  const int b = t / Lp, l = t % Lp;
  const int hh = lane >> 1;
  const size_t o = ((size_t)b * H + hh) * Lp + l;
  if (l >= L) {
    if ((lane & 1) == 0) dneg[o] = 0.f;
    return;
  }
  const size_t tok = (size_t)b * L + l;
  const float m = mtile[((size_t)b * H + hh) * (Lp / 128) + l / 128];
  const float w = exp2f(m - lse[((size_t)b * H + hh) * L + l] * 1.4426950408889634f);
  const uint4* a = reinterpret_cast<const uint4*>(y + tok * (H * 64) + lane * 32);
  const uint4* g = reinterpret_cast<const uint4*>(dy + tok * (H * 64) + lane * 32);
  uint4* gs = reinterpret_cast<uint4*>(dys + tok * (H * 64) + lane * 32);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 av = a[i], gv = g[i];
    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&av);
    const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&gv);
    uint32_t sc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g0 = __low2float(gp[k]), g1 = __high2float(gp[k]);
      acc = fmaf(__low2float(ap[k]), g0, acc);
      acc = fmaf(__high2float(ap[k]), g1, acc);
      sc[k] = pack_bf16(g0 * w, g1 * w);
    }
    gs[i] = make_uint4(sc[0], sc[1], sc[2], sc[3]);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if ((lane & 1) == 0) {
    dneg[o] = -w * acc;
    dsum[((size_t)b * H + hh) * L + l] = acc;
  }
}

// dq_acc fp32 [T, dh] (unscaled) -> bf16 dq column block of dqkv [T, 3*dh], x 1/sqrt(d)
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, size_t n8,
                                           int dh, float scale, const int* __restrict__ fallback) {
  if (*fallback != 0) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const size_t e = i * 8;
  const size_t t = e / dh, c = e % dh;
  const float4 a = reinterpret_cast<const float4*>(acc + e)[0];
  const float4 bb = reinterpret_cast<const float4*>(acc + e)[1];
  *reinterpret_cast<uint4*>(dqkv + t * (3 * (size_t)dh) + c) =
      make_uint4(pack_bf16(a.x * scale, a.y * scale), pack_bf16(a.z * scale, a.w * scale),
                 pack_bf16(bb.x * scale, bb.y * scale), pack_bf16(bb.z * scale, bb.w * scale));
}

int launch_attn_bwd_gated(const void* qkv, const void* dy, const float* lse, const float* dsum, void* dqkv, int B, int L,
                          int H, const int* gate, cudaStream_t stream);

static unsigned long long* g_fb_trace = nullptr;
static int g_fb_trace_cta = 0;
void attn_bwd_fused_set_trace(unsigned long long* buf, int cta) {
  g_fb_trace = buf;
  g_fb_trace_cta = cta;
}

// scratch layout (floats): dneg [B*H*Lp] | dsum [B*H*L] | mtile [B*H*n_t] | fallback flag (+ padding)
size_t attn_bwd_fused_stats_floats(int B, int L, int H) {
  const size_t n_t = (L + 127) / 128;
  return (size_t)B * H * (n_t * 128 + L + n_t) + 64;
}

const int* attn_bwd_fused_flag(const float* stats, int B, int L, int H) {
  const size_t n_t = (L + 127) / 128;
  return reinterpret_cast<const int*>(stats + (size_t)B * H * (n_t * 128 + L + n_t));
}

int launch_attn_bwd_fused(const void* qkv, const void* y, const void* dy, const float* lse, float* stats, float* dq_acc,
                          void* dy_scaled, void* dqkv, int B, int L, int H, int convert_dq, cudaStream_t stream) {
  OSD_CHECK(qkv && y && dy && lse && stats && dq_acc && dy_scaled && dqkv && B > 0 && L > 0 && H == 16,
            "attn_bwd_fused: bad arguments");
  const int dh = H * 64;
  const int n_t = (L + 127) / 128, Lp = n_t * 128;
  float* dneg = stats;
  float* dsum = dneg + (size_t)B * H * Lp;
  float* mtile = dsum + (size_t)B * H * L;
  int* fallback = reinterpret_cast<int*>(mtile + (size_t)B * H * n_t);
  OSD_CUDA(cudaMemsetAsync(fallback, 0, sizeof(int), stream));
  attn_bwd_tilemax_kernel<<<B * H * n_t, 128, 0, stream>>>(lse, mtile, fallback, L, n_t);
  OSD_LAUNCHED();
  attn_bwd_stats_kernel<<<ceil_div(B * Lp, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y),
                                                                static_cast<const __nv_bfloat16*>(dy), lse, mtile, dneg,
                                                                dsum, static_cast<__nv_bfloat16*>(dy_scaled), B, H, L, Lp);
  OSD_LAUNCHED();
  OSD_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * L * dh * sizeof(float), stream));
  AttnBwdFusedParams p;
  {
    uint64_t dims[3] = {(uint64_t)3 * dh, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)3 * dh * 2, (uint64_t)L * 3 * dh * 2};
    uint32_t box[3] = {64, 128, 1};
    OSD_TRY(make_tmap(&p.tma_qkv, qkv, 2, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)dh, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)dh * 2, (uint64_t)L * dh * 2};
    uint32_t box[3] = {64, 128, 1};
    OSD_TRY(make_tmap(&p.tma_dy, dy_scaled, 2, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)dh, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)dh * 4, (uint64_t)L * dh * 4};
    uint32_t box[3] = {32, 32, 1};
    OSD_TRY(make_tmap(&p.tma_dq, dq_acc, 4, 3, dims, strides, box));
  }
  p.mtile = mtile;
  p.dneg = dneg;
  p.fallback = fallback;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.B = B; p.H = H; p.L = L; p.Lp = Lp; p.dh = dh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static const int dbg_skip = [] {
    const char* e = getenv("OSD_FB_SKIP");
    return e != nullptr ? atoi(e) : 0;
  }();
  p.debug_skip = dbg_skip;
  p.trace = g_fb_trace;
  p.trace_cta = g_fb_trace_cta;
  static const bool w16 = [] {
    const char* e = getenv("OSD_FB_W16");
    return e != nullptr && e[0] == '1';
  }();
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES));
    OSD_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES));
  }
  const long long grid = (long long)n_t * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_bwd_fused: grid too large");
  if (w16)
    attn_bwd_fused_kernel<true><<<(unsigned)grid, 576, FB_SMEM_BYTES, stream>>>(p);
  else
    attn_bwd_fused_kernel<false><<<(unsigned)grid, FB_THREADS, FB_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  if (convert_dq) {
    const size_t n8 = (size_t)B * L * dh / 8;
    attn_bwd_dq_convert_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(dq_acc, p.dqkv, n8, dh, p.scale, fallback);
    OSD_LAUNCHED();
  }
  // fallback (statistics of a q tile too spread out for the w-scaling): the two-kernel path, gated on the same flag
  return launch_attn_bwd_gated(qkv, dy, lse, dsum, dqkv, B, L, H, fallback, stream);
}

}  // namespace osd
