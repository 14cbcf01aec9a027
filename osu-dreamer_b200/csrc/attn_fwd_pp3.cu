// Flash attention forward, variant "pp3" (15 / 16): the cuDNN-shaped layout.  ONE CTA per SM works on TWO 128-row q tiles of
// one (b, h) against 128-row kv tiles that both share; S and P are decoupled in TMEM; SIXTEEN softmax warps.  Kept next to the
// model's kernel ("db", variant 7) as the measured alternative: 9.2-9.5 M cycles per launch at B = 16, L = 8192 against db's
// 9.55 M (ncu), and the SAME time in a sustained loop (5.61 vs 5.55 ms) because both draw the board's 1 kW: at the cap the
// time of these kernels is their energy (profiles/r02i_fwd_power_ab.json).
//
// How it got here (profiles/r02*_fwd_*.json): every forward kernel of this repo -- two CTAs per SM with 64-row
// kv tiles ("db"), two q tiles per CTA ping-pong with P over S ("pp", removed), the same with S / P decoupled ("pp2", removed)
// -- ran at 0.76-0.84 PFLOP/s whatever its control structure, with neither the SFU (56-76 %), the tensor pipe (37 %) nor the
// issue slots (42-51 %) saturated, while cuDNN's sm100 kernel (128x128x64 tiles, 256 q rows per CTA, 16 warps) does 0.90-1.0.
// tools/micro/softmax_bench.cu isolates the cause: the per-element routine (scale FFMA2, MUFU.EX2, row-sum FADD2, bf16 pack)
// is a dependent chain, and with TWO warps per sub-partition -- 8 softmax warps per SM, which all of those kernels have -- an
// SM sustains only 11.1 elements/clk; with FOUR per sub-partition 14.7 (SFU limit: 16).  11.1 elements/clk is 2960 clk
// per 256 x 128 score block, exactly what "pp" measured (3016).  Moving exponentials to the FMA pipe adds instructions and
// makes it worse at this occupancy (10.1).  So: split every score row between two threads.
//   * softmax warp (g, h, quad): q tile g, column half h (64 of the 128 kv columns), TMEM lane quadrant quad; the two
//     partial row sums meet once, at the end, through shared memory (fixed-bound softmax: no per-step row statistic);
//   * a thread pulls its 64 scores into registers and releases the accumulator at once (s_free), so S_g(j+1) = Q_g K_{j+1}^T
//     runs under the exponentials of step j; P_g(j) goes to its own TMEM columns, so O_g += P_g(j) V_j and S_g(j+1) do not
//     order each other; the UMMA issuer is event driven (polls s_free / p_full of both groups);
//   * K / V tiles are loaded once per 256 q rows: L2 -> SM traffic 34.6 -> 17.4 GB per launch at B = 16, L = 8192.
// Fixed-bound softmax (q, k leave RMSNorm(64): per-layer score bound, no running max, no O rescale); without a finite bound
// the same kernel runs the online softmax: the two halves of a row agree on its maximum through shared memory every step
// (one named barrier per step and q tile) and half 0 rescales the row's O.
//
//   warps : 0 TMA producer | 1 TMEM allocator + UMMA issuer | 2-17 softmax: (warp - 2) = 8 h + 4 g + i, quadrant = warp % 4
//   TMEM  : S_0 [0,128) | S_1 [128,256) | P_0 bf16 [256,320) | P_1 [320,384) | O_0 [384,448) | O_1 [448,512)
//   smem  : Q_0 | Q_1 (A operands of S: TMEM is full, so S runs in SS mode) | K x3 | V x3 ([128 x 64] bf16, SW128) |
//           row-sum exchange [2][128] fp32 | barriers
#include "kernels.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace osd {

static constexpr int P3_THREADS = 576;  // 18 warps
static constexpr int P3_TILE = 128 * 128;  // bytes of a [128 x 64] bf16 tile
static constexpr int P3_STAGES = 3;
static constexpr int P3_SMEM_TILES = 2 * P3_TILE + 2 * P3_STAGES * P3_TILE;
static constexpr int P3_LSUM = (2 * 128 + 2 * 2 * 2 * 128) * 4;  // partial row sums of column half 1 per q tile | online-softmax row maxima [2 buffers][2 tiles][2 halves][128]
static constexpr int P3_SMEM_BYTES = P3_SMEM_TILES + P3_LSUM + 256 + 1024;
static constexpr uint32_t P3_TMEM_COLS = 512;

struct AttnPp3Params {
  CUtensorMap tma;  // qkv dims (3*dh, L, B), box (64, 128, 1)
  const float* bound_log2;
  __nv_bfloat16* y;
  float* lse;
  int B, H, L, dh;
  float scale_log2, scale;
};

__device__ __forceinline__ float pp3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one chunk of 32 score columns -> 32 probabilities (bf16 pairs in pk[16]); EMU of every 8 pairs use the FMA-pipe exponential
template <int EMU>
__device__ __forceinline__ void pp3_chunk(const uint32_t (&r)[32], float c, float neg_mc, uint32_t (&pk)[16], float2& s01,
                                         float2& s23) {
  const float2 c2 = make_float2(c, c), n2 = make_float2(neg_mc, neg_mc);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float2 a = ffma2(make_float2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1])), c2, n2);
    float2 e;
    if ((p & 7) < EMU)
      e = ex2_poly2(a);
    else
      e = make_float2(pp3_ex2(a.x), pp3_ex2(a.y));
    if (p & 1)
      s23 = fadd2(s23, e);
    else
      s01 = fadd2(s01, e);
    pk[p] = pack_bf16(e.x, e.y);
  }
}
// same with the columns >= valid masked to zero (last kv tile of a ragged sequence)
__device__ __forceinline__ void pp3_chunk_masked(const uint32_t (&r)[32], float c, float neg_mc, int valid, uint32_t (&pk)[16],
                                                float2& s01) {
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float e0 = (2 * p < valid) ? pp3_ex2(fmaf(__uint_as_float(r[2 * p]), c, neg_mc)) : 0.f;
    const float e1 = (2 * p + 1 < valid) ? pp3_ex2(fmaf(__uint_as_float(r[2 * p + 1]), c, neg_mc)) : 0.f;
    s01 = fadd2(s01, make_float2(e0, e1));
    pk[p] = pack_bf16(e0, e1);
  }
}

// The per-step loop of one softmax thread (q row of tile g, column half hh).  FIXED is a compile-time copy of the run-time
// "a finite score bound was given": the kernel instantiates both and branches once, so the hot fixed-bound loop carries none of
// the online path's code (interleaved, it cost 7 %: 5.33 -> 5.70 ms).
template <int EMU, bool FIXED>
__device__ __forceinline__ void pp3_softmax_loop(const AttnPp3Params& p, int n_kv, int g, int hh, int row, int lane, uint32_t tS,
                                                 uint32_t tP, uint32_t tO, uint64_t* s_full, uint64_t* s_free, uint64_t* p_full,
                                                 uint64_t* o_ready, float* sMax, float c, float& m, float& l) {
  constexpr bool fixed = FIXED;
  for (int j = 0; j < n_kv; ++j) {
    const int valid = p.L - j * 128 - hh * 64;  // valid columns of this thread's half
    mbar_wait(&s_full[g], j & 1);
    tc_fence_after();
    uint32_t rs[2][32];
    __syncwarp();
    tmem_ld32(tS, rs[0]);
    tmem_ld32(tS + 32, rs[1]);
    tmem_wait_ld();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_free[g]);  // S_g(j+1) may overwrite the accumulator from here on
    float alpha = 1.0f;
    bool o_waited = false;
    if (!fixed) {
      float mx = -INFINITY;
#pragma unroll
      for (int cch = 0; cch < 2; ++cch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cch * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(rs[cch][i]));
      float* buf = sMax + ((j & 1) * 2 + g) * 256;  // double-buffered by step parity: one barrier per step is enough
      buf[hh * 128 + row] = mx;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      mx = fmaxf(mx, buf[(hh ^ 1) * 128 + row]);
      const float m_new = fmaxf(m, mx);
      alpha = pp3_ex2((m - m_new) * c);
      m = m_new;
      if (j > 0) {
        mbar_wait(&o_ready[g], (j - 1) & 1);
        tc_fence_after();
        o_waited = true;
        if (hh == 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {  // half 0 rescales the row's O; PV_g(j) waits for its p_full arrive
#pragma unroll 1
          for (int cch = 0; cch < 4; ++cch) {  // 16 columns at a time: the 64 scores of this step stay in registers
            uint32_t ro[16];
            __syncwarp();
            tmem_ld16(tO + cch * 16, ro);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
            tmem_st16(tO + cch * 16, ro);
          }
          tmem_wait_st();
        }
      }
    }
    const float neg_mc = -m * c;
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
    // all 64 probabilities first, the P buffer second: by the time they are packed, O_g += P_g(j-1) V has long read the
    // previous P (waiting before the first store cost 9.8 % of the softmax warps' samples: the fastest warp of a group
    // reached it while the slowest was still finishing step j-1)
    uint32_t pk[2][16];
#pragma unroll
    for (int cch = 0; cch < 2; ++cch) {
      if (valid >= 64)
        pp3_chunk<EMU>(rs[cch], c, neg_mc, pk[cch], s01, s23);
      else
        pp3_chunk_masked(rs[cch], c, neg_mc, valid - cch * 32, pk[cch], s01);
    }
    if (j > 0 && !o_waited) {  // P_g's columns are free once PV_g(j-1) has read them
      mbar_wait(&o_ready[g], (j - 1) & 1);
      tc_fence_after();
    }
    __syncwarp();
    tmem_st16(tP, pk[0]);
    tmem_st16(tP + 16, pk[1]);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&p_full[g]);
    l = l * alpha + ((s01.x + s01.y) + (s23.x + s23.y));
  }
}

template <int EMU>
__global__ void __launch_bounds__(P3_THREADS, 1) attn_fwd_pp3_kernel(const __grid_constant__ AttnPp3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + P3_SMEM_TILES + P3_LSUM + 256 > dyn) __trap();
  }
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sK = sQ + 2 * P3_TILE;           // P3_STAGES tiles
  uint8_t* sV = sK + P3_STAGES * P3_TILE;   // P3_STAGES tiles
  float* sL = reinterpret_cast<float*>(sV + P3_STAGES * P3_TILE);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + P3_STAGES * P3_TILE + P3_LSUM);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2] S_g(j) complete
  uint64_t* s_free = bars + 15;   // [2] S_g(j) is in the registers of group g's 4 warps: the accumulator may be overwritten
  uint64_t* p_full = bars + 17;   // [2] P_g(j) written by the 4 warps of group g
  uint64_t* o_ready = bars + 19;  // [2] O_g += P_g(j) V_j complete: P_g's columns may be overwritten (and O rescaled)
  uint64_t* acc_done = bars + 21; // [2] last PV of group g complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qp = (p.L + 255) / 256;  // q-tile pairs
  const int qp = blockIdx.x % n_qp;
  const int bh = blockIdx.x / n_qp;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qp * 256;
  const int n_kv = (p.L + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma);
    mbar_init(q_full, 1);
    for (int i = 0; i < P3_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 8);
      mbar_init(&p_full[g], 8);
      mbar_init(&o_ready[g], 1);
      mbar_init(&acc_done[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, P3_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * P3_TILE);
      tma_load_3d(sQ, &p.tma, q_full, h * 64, q0, b);
      tma_load_3d(sQ + P3_TILE, &p.tma, q_full, h * 64, q0 + 128, b);  // rows past L are zero-filled by TMA
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], P3_TILE);
        tma_load_3d(sK + st * P3_TILE, &p.tma, &k_full[st], p.dh + h * 64, j * 128, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], P3_TILE);
        tma_load_3d(sV + st * P3_TILE, &p.tma, &v_full[st], 2 * p.dh + h * 64, j * 128, b);
        if (++st == P3_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== UMMA issuer (event driven)
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(FMT_BF16, 0, 0, 128, 128);  // S = Q K^T : A, B K-major in smem, N = 128 kv
      const uint32_t idesc_o = make_idesc(FMT_BF16, 0, 1, 128, 64);   // O += P V : A in TMEM, B MN-major, N = 64 d
      auto issue_s = [&](int g, int j) {
        const uint32_t aQ = smem_u32(sQ + g * P3_TILE), aK = smem_u32(sK + (j % P3_STAGES) * P3_TILE);
        const uint32_t tS = tmem_base + g * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tS, make_smem_desc(aQ + k * 32, 0, 1024), make_smem_desc(aK + k * 32, 0, 1024), idesc_s, k > 0);
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int j) {
        const uint32_t aV = smem_u32(sV + (j % P3_STAGES) * P3_TILE);
        const uint32_t tP = tmem_base + 256 + g * 64, tO = tmem_base + 384 + g * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // contraction over the 128 kv rows, 16 per instruction
          umma_f16_ts(tO, tP + k * 8, make_smem_desc(aV + k * 16 * 128, 0, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&o_ready[g]);
        if (j + 1 == n_kv) umma_commit(&acc_done[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(&k_empty[0]);
      int js[2] = {1, 1};  // next S tile to issue per group
      int jp[2] = {0, 0};  // next PV tile to issue per group
      const long long t_start = clock64();
      while (jp[0] < n_kv || jp[1] < n_kv) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int j = js[g];
          // S_g(j): the group has S_g(j-1) in registers and K_j has landed
          if (j < n_kv && mbar_test(&s_free[g], (j - 1) & 1) && mbar_test(&k_full[j % P3_STAGES], (j / P3_STAGES) & 1)) {
            tc_fence_after();
            issue_s(g, j);
            if (js[g ^ 1] > j) umma_commit(&k_empty[j % P3_STAGES]);  // the other group has used K_j already
            js[g] = j + 1;
            progress = true;
          }
          const int i = jp[g];
          // O_g += P_g(i) V_i: the group has published P_g(i) and V_i has landed
          if (i < n_kv && mbar_test(&p_full[g], i & 1) && mbar_test(&v_full[i % P3_STAGES], (i / P3_STAGES) & 1)) {
            tc_fence_after();
            issue_pv(g, i);
            if (jp[g ^ 1] > i) umma_commit(&v_empty[i % P3_STAGES]);
            jp[g] = i + 1;
            progress = true;
          }
        }
        if (!progress) {
          __nanosleep(20);
          if (clock64() - t_start > OSD_WATCHDOG_CYCLES) mbar_timeout(&p_full[0], 0);
        }
      }
    }
  } else {
    // ================================================================== softmax (thread = q row of tile g, column half hh)
    const int widx = warp - 2;
    const int g = (widx >> 2) & 1, hh = widx >> 3;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + hh * 64 + lane_off;
    const uint32_t tP = tmem_base + 256 + g * 64 + hh * 32 + lane_off;
    const uint32_t tO = tmem_base + 384 + g * 64 + lane_off;
    const float c = p.scale_log2;
    float bound = INFINITY;
    if (p.bound_log2 != nullptr) bound = __ldg(p.bound_log2);
    const bool fixed = bound < 3.0e38f;  // else: online softmax, the two halves of a row agree on its maximum every step
    float m = fixed ? bound / c : -INFINITY, l = 0.f;
    float* sMax = sL + 2 * 128;
    if (fixed)
      pp3_softmax_loop<EMU, true>(p, n_kv, g, hh, row, lane, tS, tP, tO, s_full, s_free, p_full, o_ready, sMax, c, m, l);
    else
      pp3_softmax_loop<EMU, false>(p, n_kv, g, hh, row, lane, tS, tP, tO, s_full, s_free, p_full, o_ready, sMax, c, m, l);
    // the two column halves of a row meet here: half 1 publishes its partial sum, half 0 finishes the row
    if (hh == 1) sL[g * 128 + row] = l;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");  // the 8 warps of q tile g
    if (hh == 0) {
      l += sL[g * 128 + row];
      mbar_wait(&acc_done[g], 0);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int q = q0 + g * 128 + row;
      const bool ok = q < p.L;
#pragma unroll 1
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tO + cch * 32, r);
        tmem_wait_ld();
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.y + ((size_t)b * p.L + q) * p.dh + h * 64 + cch * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(pack_bf16(__uint_as_float(r[8 * i]) * inv_l, __uint_as_float(r[8 * i + 1]) * inv_l),
                                pack_bf16(__uint_as_float(r[8 * i + 2]) * inv_l, __uint_as_float(r[8 * i + 3]) * inv_l),
                                pack_bf16(__uint_as_float(r[8 * i + 4]) * inv_l, __uint_as_float(r[8 * i + 5]) * inv_l),
                                pack_bf16(__uint_as_float(r[8 * i + 6]) * inv_l, __uint_as_float(r[8 * i + 7]) * inv_l));
        }
      }
      if (ok && p.lse != nullptr) p.lse[((size_t)b * p.H + h) * p.L + q] = m * p.scale + __logf(l);
    }
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, P3_TMEM_COLS);
  }
}

template <int EMU>
static int launch_pp3_t(const AttnPp3Params& p, long long grid, cudaStream_t stream) {
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(attn_fwd_pp3_kernel<EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM_BYTES));
  }
  attn_fwd_pp3_kernel<EMU><<<(unsigned)grid, P3_THREADS, P3_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  return 0;
}

// emu: share of the exponentials on the FMA pipe, in EIGHTHS (0 .. 4); < 0 = the default (OSD_PP_EMU or 2)
int launch_attn_fwd_pp3(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H, int emu,
                       cudaStream_t stream) {
  OSD_CHECK(qkv && y && B > 0 && L > 0 && H > 0, "attn_fwd_pp3: bad arguments");
  AttnPp3Params p;
  const int dh = H * 64;
  uint64_t dims[3] = {(uint64_t)3 * dh, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)3 * dh * 2, (uint64_t)L * 3 * dh * 2};
  uint32_t box[3] = {64, 128, 1};
  OSD_TRY(make_tmap(&p.tma, qkv, 2, 3, dims, strides, box));
  p.bound_log2 = bound_log2;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.lse = lse;
  p.B = B; p.H = H; p.L = L; p.dh = dh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static const int emu_default = [] {
    const char* e = getenv("OSD_PP_EMU");
    return (e != nullptr && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 2;
  }();
  if (emu < 0) emu = emu_default;
  const long long grid = (long long)ceil_div(L, 256) * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_fwd_pp3: grid too large");
  {
    if (emu == 1)
      OSD_TRY(launch_pp3_t<1>(p, grid, stream));
    else if (emu == 2)
      OSD_TRY(launch_pp3_t<2>(p, grid, stream));
    else if (emu == 3)
      OSD_TRY(launch_pp3_t<3>(p, grid, stream));
    else if (emu == 4)
      OSD_TRY(launch_pp3_t<4>(p, grid, stream));
    else
      OSD_TRY(launch_pp3_t<0>(p, grid, stream));
  }
  return 0;
}

}  // namespace osd
