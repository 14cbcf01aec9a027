// Flash attention forward, variant "db": two S accumulators in TMEM ping-pong, so the QK^T MMA of tile j+1 (and
// j+2) runs while the softmax warps are still working on tile j -- the softmax never waits for the tensor pipe
// and the SFU (the binding unit at head_dim 64) stays busy.  Profile that motivated it (ncu, variant 2):
// 20 % of all warp samples sat on the s_full wait because the co-resident CTAs fall into lockstep.
//   CTA = 128 q rows of one (b,h); 64-row kv tiles, 3-stage K/V ring; 2 CTAs/SM (192 of 256 TMEM columns)
//   TMEM: S0 cols [0,64) | S1 [64,128) | O [128,192); bf16 P_j is written back over the first 32 columns of
//         S_{j&1} and consumed from there as the A operand of O += P_j V_j; S_{j+2} is issued right after that MMA
//         (in-order tensor pipe), so the aliasing is safe.
#include "kernels.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace osd {

#ifndef OSD_DB_EMU
#define OSD_DB_EMU 0
#endif
static constexpr bool DB_EMU = OSD_DB_EMU != 0;

static constexpr int DB_THREADS = 192;
static constexpr int DB_T128 = 128 * 128;
static constexpr int DB_T64 = 64 * 128;
static constexpr int DB_STAGES = 3;
static constexpr int DB_SMEM_TILES = DB_T128 + 2 * DB_STAGES * DB_T64;
static constexpr int DB_ONES = 2048;  // variant 8: [16 x 64] bf16 tile of ones (B operand of the row-sum MMA)
static constexpr int DB_SMEM_BYTES = DB_SMEM_TILES + DB_ONES + 256;
// X3 (fp32-grade) layout: Q_hi | Q_lo staging, K x3, (V_hi | V_lo) x3
static constexpr int DB_SMEM_TILES_X3 = 2 * DB_T128 + DB_STAGES * DB_T64 + 2 * DB_STAGES * DB_T64;
static constexpr int DB_SMEM_BYTES_X3 = DB_SMEM_TILES_X3 + DB_ONES + 256;
static constexpr uint32_t DB_TMEM_COLS = 256;

struct AttnDbParams {
  CUtensorMap tma_q;   // dims (3*dh, L, B), box (64, 128, 1)
  CUtensorMap tma_kv;  // box (64, 64, 1)
  const float* bound_log2;
  __nv_bfloat16* y;
  float* lse;
  int B, H, L, dh;
  float scale_log2, scale;
  int only_if_online;  // 1: return immediately when the bound is finite (another kernel has done the work)
  int lo_col;          // X3: first column of the lo block of qkv (= 3 * dh)
};

__device__ __forceinline__ float db_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// PF (variant 6): in fixed-bound mode the softmax warps (a) probe the barriers a step needs (o_ready of the previous
// PV, s_full of the next S) with non-blocking test_waits BEFORE the exponentials, so the probe latency hides behind
// the math, and (b) issue the TMEM load of S_{j+1} before storing / publishing P_j, so the load latency hides behind
// the store + fence + arrive tail.  ncu source view of the non-PF kernel: 10 % of the softmax warps' samples sit on
// the s_full probe, 5 % on the o_ready probe, 7 % on the TMEM load and 9 % on the store tail.
// DR (variant 8, needs QT; active when the bound is <= 60 octaves, else the kernel behaves like variant 7): fewer
// instructions per score in the softmax warps, because the step is power-bound:
//  * Q is multiplied by scale*log2(e) (and re-rounded to bf16) while it is copied into TMEM, so S arrives in octaves and
//    P = 2^S needs no FFMA; without the "- bound" shift 2^S <= 2^60 and the 8192-term row sums stay far inside fp32 / bf16
//    range, and the shift cancels in O / l anyway;
//  * the row sum l = sum_j P_ij is accumulated by the tensor core: one extra N=16 MMA per kv step, P (TMEM) x a tile of
//    ones (shared memory) into 16 more TMEM columns -- 1/8 of the PV MMA's time on a pipe that is 38 % busy -- instead
//    of 32 FADD2 per thread and step; it sums the bf16-rounded P, the same values the PV MMA multiplies.
// QT (variant 7): the Q tile is copied into TMEM once (columns [192, 224)) and is the A operand of S = Q K^T from
// there (tcgen05.mma .ts), which halves the shared-memory operand reads of the kernel (Q 16 KB + K 8 KB + V 8 KB per
// kv step -> 16 KB).
// X3 (needs QT, excludes PF / DR): the fp32-grade attention of precision='fp32' (BASELINE configs[2]) on this kernel's
// pipeline.  qkv arrives as bf16 (hi | lo) pairs -- [T, 2*3*dh] = hi block | lo block, each (q | k | v) -- and the two
// products run as  S = Q_hi K_hi^T + Q_lo K_hi^T  and  O += P_hi V_hi + P_hi V_lo : the correction terms that carry the
// rounding of the operand that is NOT averaged over (q for a score row, v for an output row), four MMA units instead of the
// six of the full three-term split.  tools/x3_terms_sweep.py (profiles/r02j_x3_terms_sweep.json) measured every subset on
// the single-buffered kernel this replaces (attn_fwd_x3.cu): this one keeps one forward at 8.7e-5 and the 64-step sampler at
// 4.0e-5 of the fp32 reference (tolerance 1e-3; all six units: 1.7e-5 / 2.7e-5; plain bf16 attention: 3.3e-4 / 1.4e-4).  No
// P_lo and no K_lo exist, so the softmax is the bf16 kernel's except that the row sum is taken over the ROUNDED P (what
// the MMA multiplies: a row dominated by one key stays exact).  TMEM: ... | Q_hi [192,224) | Q_lo [224,256).  y leaves
// as (hi | lo) pairs [T, 2*dh].
// MC (variant 17, needs QT): the CTAs of two adjacent q tiles of one (b, h) form a thread-block cluster and share every K / V
// tile: each loads HALF of it (32 of the 64 kv rows) and TMA-multicasts that half into both CTAs' shared memory, so the
// L2 -> SM traffic of the kernel halves (34.6 -> 17.3 GB per launch at B = 16, L = 8192) -- the kernel is energy bound
// (DESIGN.md 5b), bytes are time.  A stage is refilled only when BOTH CTAs' MMAs have read it: the empty barriers count two
// arrivals and the tcgen05.commit that releases a stage is multicast to the pair.
template <bool PF, bool QT, bool DR = false, bool X3 = false, bool MC = false>
__global__ void __launch_bounds__(DB_THREADS, 2) attn_fwd_db_kernel(const __grid_constant__ AttnDbParams p) {
  static_assert(!X3 || (QT && !PF && !DR), "X3 runs on the QT pipeline only");
  static_assert(!MC || (QT && !PF && !DR && !X3), "MC runs on the plain QT pipeline only");
  extern __shared__ uint8_t smem_raw[];
  if (p.only_if_online && p.bound_log2 != nullptr && *p.bound_log2 < 3.0e38f) return;
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + (X3 ? DB_SMEM_TILES_X3 : DB_SMEM_TILES) + DB_ONES + 160 > dyn) __trap();
  }
  constexpr int VT = X3 ? 2 * DB_T64 : DB_T64;  // bytes of one V stage (X3: V_hi | V_lo)
  uint8_t* sQ = smem;                                // X3: Q_hi | Q_lo
  uint8_t* sK = sQ + (X3 ? 2 : 1) * DB_T128;
  uint8_t* sV = sK + DB_STAGES * DB_T64;
  uint8_t* sOnes = sV + DB_STAGES * VT;  // 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + DB_ONES);
  // direct mode (DR): decided once per CTA from the bound
  bool direct = false;
  if (DR) {
    const float bnd = p.bound_log2 != nullptr ? *p.bound_log2 : INFINITY;
    direct = bnd <= 60.0f;
  }
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2]
  uint64_t* p_full = bars + 15;
  uint64_t* o_ready = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  uint64_t* qt_ready = bars + 18;  // QT: Q copied into TMEM by the softmax warps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = MC ? ((p.L + 255) / 256) * 2 : (p.L + 127) / 128;  // MC: padded to whole pairs (a pad CTA computes, writes nothing)
  const int qt = blockIdx.x % n_qt;
  const int bh = blockIdx.x / n_qt;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * 128;
  const int n_kv = (p.L + 63) / 64;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < DB_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], MC ? 2 : 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], MC ? 2 : 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 4);
    mbar_init(o_ready, 1);
    mbar_init(qt_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, DB_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, (X3 ? 2 : 1) * DB_T128);
      tma_load_3d(sQ, &p.tma_q, q_full, h * 64, q0, b);
      if (X3) tma_load_3d(sQ + DB_T128, &p.tma_q, q_full, p.lo_col + h * 64, q0, b);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], DB_T64);
        if (MC)  // my half of the tile (rows 32 crank .. +32) into both CTAs; the peer sends the other half
          tma_load_3d_multicast(sK + st * DB_T64 + crank * (DB_T64 / 2), &p.tma_kv, &k_full[st], p.dh + h * 64,
                                j * 64 + (int)crank * 32, b, 0x3);
        else
          tma_load_3d(sK + st * DB_T64, &p.tma_kv, &k_full[st], p.dh + h * 64, j * 64, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], VT);
        if (MC)
          tma_load_3d_multicast(sV + st * VT + crank * (DB_T64 / 2), &p.tma_kv, &v_full[st], 2 * p.dh + h * 64,
                                j * 64 + (int)crank * 32, b, 0x3);
        else
          tma_load_3d(sV + st * VT, &p.tma_kv, &v_full[st], 2 * p.dh + h * 64, j * 64, b);
        if (X3) tma_load_3d(sV + st * VT + DB_T64, &p.tma_kv, &v_full[st], p.lo_col + 2 * p.dh + h * 64, j * 64, b);
        if (++st == DB_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(FMT_BF16, 0, 0, 128, 64);
      const uint32_t idesc_o = make_idesc(FMT_BF16, 0, 1, 128, 64);
      const uint32_t idesc_l = make_idesc(FMT_BF16, 0, 0, 128, 16);
      const uint32_t aQ = smem_u32(sQ);
      const uint32_t tO = tmem_base + 128;
      auto issue_s = [&](int j) {
        const int st = j % DB_STAGES;
        mbar_wait(&k_full[st], (j / DB_STAGES) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * DB_T64);
        const uint32_t tS = tmem_base + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (QT)
            umma_f16_ts(tS, tmem_base + 192 + k * 8, make_smem_desc(aK + k * 32, 0, 1024), idesc_s, k > 0);
          else
            umma_f16_ss(tS, make_smem_desc(aQ + k * 32, 0, 1024), make_smem_desc(aK + k * 32, 0, 1024), idesc_s, k > 0);
        }
        if (X3) {  // + Q_lo K_hi^T
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tS, tmem_base + 224 + k * 8, make_smem_desc(aK + k * 32, 0, 1024), idesc_s, 1u);
        }
        if (MC)
          umma_commit_multicast(&k_empty[st], 0x3);
        else
          umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      };
      if (QT) {
        mbar_wait(qt_ready, 0);
        tc_fence_after();
      } else {
        mbar_wait(q_full, 0);
      }
      issue_s(0);
      if (n_kv > 1) issue_s(1);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % DB_STAGES;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j / DB_STAGES) & 1);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + st * VT);
        const uint32_t tP = tmem_base + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tO, tP + k * 8, make_smem_desc(aV + k * 16 * 128, 0, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        if (X3) {  // + P_hi V_lo
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tO, tP + k * 8, make_smem_desc(aV + DB_T64 + k * 16 * 128, 0, 1024), idesc_o, 1u);
        }
        if (DR && direct) {  // l += P_j x ones  (columns [224, 240))
          const uint32_t aOnes = smem_u32(sOnes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tmem_base + 224, tP + k * 8, make_smem_desc(aOnes + k * 32, 0, 1024), idesc_l,
                        (j > 0 || k > 0) ? 1u : 0u);
        }
        if (MC)
          umma_commit_multicast(&v_empty[st], 0x3);
        else
          umma_commit(&v_empty[st]);
        umma_commit(o_ready);
        if (j + 2 < n_kv) issue_s(j + 2);  // reuses S_{j&1}: issued after the MMA that read P_j from it
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tO = tmem_base + 128 + lane_off;
    if (QT) {  // one-time copy of this thread's Q row (128 B, SW128 smem) into TMEM
      mbar_wait(q_full, 0);
      const uint32_t base = smem_u32(sQ) + row * 128;
      uint32_t rq[32];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rq[4 * u]), "=r"(rq[4 * u + 1]), "=r"(rq[4 * u + 2]), "=r"(rq[4 * u + 3])
                     : "r"(base + ((u ^ (row & 7)) << 4)));
      if (DR && direct) {
        const float cq = p.scale_log2;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float lo = __uint_as_float(rq[u] << 16) * cq, hi = __uint_as_float(rq[u] & 0xffff0000u) * cq;
          rq[u] = pack_bf16(lo, hi);
        }
        // the tile of ones: 2048 bytes, 16 per thread of the 128 softmax threads
        *reinterpret_cast<uint4*>(sOnes + (threadIdx.x - 64) * 16) = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        fence_proxy_async_smem();
      }
      __syncwarp();
      tmem_st32(tmem_base + 192 + lane_off, rq);
      if (X3) {  // the same for this thread's Q_lo row
        const uint32_t base_lo = base + DB_T128;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rq[4 * u]), "=r"(rq[4 * u + 1]), "=r"(rq[4 * u + 2]), "=r"(rq[4 * u + 3])
                       : "r"(base_lo + ((u ^ (row & 7)) << 4)));
        __syncwarp();
        tmem_st32(tmem_base + 224 + lane_off, rq);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qt_ready);
    }
    const float c = (DR && direct) ? 1.0f : p.scale_log2;  // direct mode: Q was pre-scaled, S is already in octaves
    float bound = INFINITY;
    if (p.bound_log2 != nullptr) bound = __ldg(p.bound_log2);
    const bool fixed = bound < 3.0e38f;
    float m = (DR && direct) ? 0.f : (fixed ? bound / c : -INFINITY), l = 0.f;
    const bool pf = PF && fixed;
    uint32_t r0[32], r1[32];
    bool have = false;  // r0 / r1 already hold (or are receiving) S_j
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t tS = tmem_base + (j & 1) * 64 + lane_off;
      if (!have) {
        mbar_wait(&s_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        __syncwarp();
        tmem_ld32(tS, r0);
        tmem_ld32(tS + 32, r1);
      }
      bool o_ok = false, s_ok = false;
      if (pf) {
        if (j > 0) o_ok = mbar_test(o_ready, (j - 1) & 1);
        if (j + 1 < n_kv) s_ok = mbar_test(&s_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
      }
      const int valid = p.L - j * 64;
      tmem_wait_ld();
      float m_new = m, alpha = 1.0f;
      if (!fixed) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < valid) mx = fmaxf(mx, __uint_as_float(r0[i]));
          if (32 + i < valid) mx = fmaxf(mx, __uint_as_float(r1[i]));
        }
        m_new = fmaxf(m, mx);
        alpha = db_ex2((m - m_new) * c);
      }
      const float neg_mc = -m_new * c;
      const float2 c2 = make_float2(c, c), n2 = make_float2(neg_mc, neg_mc);
      uint32_t pk[32];
      float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
      if (DR && direct && valid >= 64) {  // P = 2^S: no shift, no row sum (the tensor core accumulates it)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          pk[i >> 1] = pack_bf16(db_ex2(__uint_as_float(r0[i])), db_ex2(__uint_as_float(r0[i + 1])));
          pk[16 + (i >> 1)] = pack_bf16(db_ex2(__uint_as_float(r1[i])), db_ex2(__uint_as_float(r1[i + 1])));
        }
      } else if (valid >= 64) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 a = ffma2(make_float2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])), c2, n2);
          const float2 bb = ffma2(make_float2(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])), c2, n2);
          const float2 ea = make_float2(db_ex2(a.x), db_ex2(a.y));
          // half of the exponentials on the FMA pipe (ex2_poly2): the SFU (16 ex2 / clk / SM) is the binding unit at d = 64
          const float2 eb = DB_EMU ? ex2_poly2(bb) : make_float2(db_ex2(bb.x), db_ex2(bb.y));
          pk[i >> 1] = pack_bf16(ea.x, ea.y);
          pk[16 + (i >> 1)] = pack_bf16(eb.x, eb.y);
          if (X3) {  // normalise by what the MMA multiplies: the bf16-rounded probabilities
            s01 = fadd2(s01, make_float2(__uint_as_float(pk[i >> 1] << 16), __uint_as_float(pk[i >> 1] & 0xffff0000u)));
            s23 = fadd2(s23, make_float2(__uint_as_float(pk[16 + (i >> 1)] << 16), __uint_as_float(pk[16 + (i >> 1)] & 0xffff0000u)));
          } else {
            s01 = fadd2(s01, ea);
            s23 = fadd2(s23, eb);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = (i < valid) ? db_ex2(fmaf(__uint_as_float(r0[i]), c, neg_mc)) : 0.f;
          const float a1 = (i + 1 < valid) ? db_ex2(fmaf(__uint_as_float(r0[i + 1]), c, neg_mc)) : 0.f;
          const float b0 = (32 + i < valid) ? db_ex2(fmaf(__uint_as_float(r1[i]), c, neg_mc)) : 0.f;
          const float b1 = (33 + i < valid) ? db_ex2(fmaf(__uint_as_float(r1[i + 1]), c, neg_mc)) : 0.f;
          pk[i >> 1] = pack_bf16(a0, a1);
          pk[16 + (i >> 1)] = pack_bf16(b0, b1);
          if (X3) {
            s01 = fadd2(s01, make_float2(__uint_as_float(pk[i >> 1] << 16), __uint_as_float(pk[i >> 1] & 0xffff0000u)));
            s23 = fadd2(s23, make_float2(__uint_as_float(pk[16 + (i >> 1)] << 16), __uint_as_float(pk[16 + (i >> 1)] & 0xffff0000u)));
          } else {
            s01 = fadd2(s01, make_float2(a0, a1));
            s23 = fadd2(s23, make_float2(b0, b1));
          }
        }
      }
      const float sum = (s01.x + s01.y) + (s23.x + s23.y);
      // O_{j-1} must be complete before (a) it is rescaled (online mode) and (b) this warp runs further ahead of the
      // o_ready phase counter; by now that MMA has long retired, so this wait is free in steady state.
      if (j > 0) {
        if (!o_ok) mbar_wait(o_ready, (j - 1) & 1);
        tc_fence_after();
        if (!fixed && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int cch = 0; cch < 2; ++cch) {
            uint32_t ro[32];
            __syncwarp();
            tmem_ld32(tO + cch * 32, ro);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
            tmem_st32(tO + cch * 32, ro);
          }
          tmem_wait_st();
        }
      }
      __syncwarp();
      have = false;
      if (pf && j + 1 < n_kv) {
        // r0 / r1 are dead (pk holds P_j): S_{j+1} streams in from the other accumulator while P_j is stored / published
        if (!s_ok) mbar_wait(&s_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t tSn = tmem_base + ((j + 1) & 1) * 64 + lane_off;
        tmem_ld32(tSn, r0);
        tmem_ld32(tSn + 32, r1);
        have = true;
      }
      tmem_st32(tS, pk);  // P_j (64 bf16 = 32 columns) over the consumed S_j
      tmem_wait_st();
      l = l * alpha + sum;
      m = m_new;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_ready, (n_kv - 1) & 1);
    tc_fence_after();
    if (DR && direct) {  // the row sums accumulated by the ones-tile MMA (all 16 columns are equal)
      uint32_t rl[16];
      __syncwarp();
      tmem_ld16(tmem_base + 224 + lane_off, rl);
      tmem_wait_ld();
      l = __uint_as_float(rl[0]);
    }
    const float inv_l = 1.0f / l;
    const int q = q0 + row;
    const bool ok = q < p.L;
#pragma unroll 1
    for (int cch = 0; cch < 2; ++cch) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tO + cch * 32, r);
      tmem_wait_ld();
      if (ok && !X3) {
        uint4* dst = reinterpret_cast<uint4*>(p.y + ((size_t)b * p.L + q) * p.dh + h * 64 + cch * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_uint4(pack_bf16(__uint_as_float(r[8 * i]) * inv_l, __uint_as_float(r[8 * i + 1]) * inv_l),
                              pack_bf16(__uint_as_float(r[8 * i + 2]) * inv_l, __uint_as_float(r[8 * i + 3]) * inv_l),
                              pack_bf16(__uint_as_float(r[8 * i + 4]) * inv_l, __uint_as_float(r[8 * i + 5]) * inv_l),
                              pack_bf16(__uint_as_float(r[8 * i + 6]) * inv_l, __uint_as_float(r[8 * i + 7]) * inv_l));
      }
      if (ok && X3) {  // y as (hi | lo) pairs: [T, 2 * dh]
        __nv_bfloat16* yrow = p.y + ((size_t)b * p.L + q) * (2 * p.dh) + h * 64 + cch * 32;
        uint4* dh4 = reinterpret_cast<uint4*>(yrow);
        uint4* dl4 = reinterpret_cast<uint4*>(yrow + p.dh);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = __uint_as_float(r[8 * i + 2 * e]) * inv_l, v1 = __uint_as_float(r[8 * i + 2 * e + 1]) * inv_l;
            hw[e] = pack_bf16(v0, v1);
            lw[e] = pack_bf16(v0 - __uint_as_float(hw[e] << 16), v1 - __uint_as_float(hw[e] & 0xffff0000u));
          }
          dh4[i] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          dl4[i] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    }
    if (ok && p.lse != nullptr) p.lse[((size_t)b * p.H + h) * p.L + q] = m * p.scale + __logf(l);
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, DB_TMEM_COLS);
  }
  if (MC) cluster_sync_all();  // neither CTA leaves while the other may still multicast into it or arrive on its barriers
}

int launch_attn_fwd_db_gated(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                             int only_if_online, cudaStream_t stream);
template <bool PF, bool QT, bool DR = false, bool X3 = false, bool MC = false>
static int launch_attn_fwd_db_t(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                                int only_if_online, cudaStream_t stream);
// fp32-grade attention (precision='fp32'): qkv bf16 [T, 2*3*dh] (hi block | lo block), y bf16 [T, 2*dh] (hi | lo)
int launch_attn_fwd_db_x3(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream) {
  return launch_attn_fwd_db_t<false, true, false, true>(qkv, y, lse, bound_log2, B, L, H, 0, stream);
}
int launch_attn_fwd_db(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                       cudaStream_t stream) {
  return launch_attn_fwd_db_t<false, false>(qkv, y, lse, bound_log2, B, L, H, 0, stream);
}
int launch_attn_fwd_db_qt(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream) {
  return launch_attn_fwd_db_t<false, true>(qkv, y, lse, bound_log2, B, L, H, 0, stream);
}
int launch_attn_fwd_db_dr(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream) {
  return launch_attn_fwd_db_t<false, true, true>(qkv, y, lse, bound_log2, B, L, H, 0, stream);
}
int launch_attn_fwd_db_pf(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream) {
  return launch_attn_fwd_db_t<true, false>(qkv, y, lse, bound_log2, B, L, H, 0, stream);
}
int launch_attn_fwd_db_gated(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                             int only_if_online, cudaStream_t stream) {
  return launch_attn_fwd_db_t<false, false>(qkv, y, lse, bound_log2, B, L, H, only_if_online, stream);
}
// variant 17: K / V tiles shared by a 2-CTA cluster through TMA multicast
int launch_attn_fwd_db_mc(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream) {
  return launch_attn_fwd_db_t<false, true, false, false, true>(qkv, y, lse, bound_log2, B, L, H, 0, stream);
}
template <bool PF, bool QT, bool DR, bool X3, bool MC>
static int launch_attn_fwd_db_t(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                                int only_if_online, cudaStream_t stream) {
  OSD_CHECK(qkv && y && B > 0 && L > 0 && H > 0, "attn_fwd_db: bad arguments");
  AttnDbParams p;
  const int dh = H * 64;
  const uint64_t width = (uint64_t)(X3 ? 6 : 3) * dh;  // X3: (hi block | lo block)
  uint64_t dims[3] = {width, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {width * 2, (uint64_t)L * width * 2};
  uint32_t box_q[3] = {64, 128, 1}, box_kv[3] = {64, MC ? 32u : 64u, 1};  // MC: each CTA of the pair loads half a kv tile
  OSD_TRY(make_tmap(&p.tma_q, qkv, 2, 3, dims, strides, box_q));
  OSD_TRY(make_tmap(&p.tma_kv, qkv, 2, 3, dims, strides, box_kv));
  p.bound_log2 = bound_log2;
  p.lo_col = 3 * dh;
  p.only_if_online = only_if_online;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.lse = lse;
  p.B = B; p.H = H; p.L = L; p.dh = dh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(attn_fwd_db_kernel<PF, QT, DR, X3, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  X3 ? DB_SMEM_BYTES_X3 : DB_SMEM_BYTES));
  }
  const long long grid = (long long)(MC ? 2 * ceil_div(L, 256) : ceil_div(L, 128)) * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_fwd_db: grid too large");
  if (MC) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(DB_THREADS);
    cfg.dynamicSmemBytes = DB_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OSD_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_db_kernel<PF, QT, DR, X3, MC>, p));
    OSD_LAUNCHED();
    return 0;
  }
  attn_fwd_db_kernel<PF, QT, DR, X3, MC><<<(unsigned)grid, DB_THREADS, X3 ? DB_SMEM_BYTES_X3 : DB_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
