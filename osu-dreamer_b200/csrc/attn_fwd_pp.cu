// Flash attention forward, variant "pp" (ping-pong): ONE CTA per SM works on TWO 128-row q tiles of one (b, h) against
// 128-row kv tiles; each q tile has its own softmax warpgroup (4 warps, thread = q row) and its own S / O accumulators in
// TMEM, and both share every K / V tile that TMA brings in.
//
// Why (ncu of the library kernel to beat, profiles/r02c_cudnn_sdpa_summary.json: cuDNN's sm100 flash fprop runs 128x128x64
// tiles, 256 q rows per CTA, 1 CTA/SM and reaches 1.0 PFLOP/s at d = 64 where the 2-CTA/SM, 64-row-kv "db" kernel does 0.82):
//  * the binding unit at d = 64 is the SFU (16 ex2 / clk / SM).  A softmax warp alternates between ~64 MUFU per 64 score
//    columns and a fixed ~400 clk off the SFU per kv step (TMEM load latency, pack, TMEM store, fences, barrier round trip).
//    128-row kv tiles halve the number of steps, so the fixed part is paid half as often per exponential.
//  * the two warpgroups ping-pong by construction: after a group publishes P_j it cannot continue until the tensor pipe has
//    run O += P_j V_j and S_{j+1} = Q K_{j+1}^T for its tile (512 clk) -- exactly when the other group has the SFU to itself.
//    Two independent CTAs per SM have no such coupling and drift through all phase offsets (an explicit start offset
//    changed nothing: profiles/r02b_fwd_stagger_ab.json).
//  * K / V tiles are loaded once for 256 q rows: L2 -> SM traffic 34.6 GB -> 17 GB per launch at B = 16, L = 8192.
//  * a share of the exponentials runs on the FMA pipe (ex2_poly2, degree-3 minimax, 7.5e-5 before the bf16 rounding): with
//    the SFU saturated this now shortens the step (it did not in the db kernel, whose SFU was 76 % busy with stalls
//    elsewhere); cuDNN's XU-pipe count shows it emulates ~24 % of its exponentials as well.
//
//   warps : 0 TMA producer | 1 TMEM allocator + UMMA issuer | 2-5 softmax group 0 (q tile 0) | 6-9 softmax group 1
//   TMEM  : S_0 [0,128) | S_1 [128,256) | O_0 [256,320) | O_1 [320,384) | Q_0 bf16 [384,416) | Q_1 [416,448)
//           bf16 P_j is written back over the first 64 columns of S_g as the exponentials are produced (chunk c of 32
//           scores -> columns [16c, 16c+16), always behind the load front) and is the TMEM A operand of O_g += P_j V_j;
//           S_g(j+1) is issued after that MMA (in-order tensor pipe).  Q lives in TMEM as the A operand of S (as in "db").
//   smem  : Q_0 | Q_1 (staging for the TMEM copy) | K x3 | V x3 ([128 x 64] bf16, SWIZZLE_128B) | barriers
// Fixed-bound softmax (q, k leave RMSNorm(64): per-layer score bound, no running max, no O rescale); when no finite bound
// is given the same kernel runs the online softmax (two TMEM passes per step + O rescale by the softmax warps).
#include "kernels.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace osd {

static constexpr int PP_THREADS = 320;
static constexpr int PP_TILE = 128 * 128;  // bytes of a [128 x 64] bf16 tile
static constexpr int PP_STAGES = 3;
static constexpr int PP_SMEM_TILES = 2 * PP_TILE + 2 * PP_STAGES * PP_TILE;
static constexpr int PP_SMEM_BYTES = PP_SMEM_TILES + 256 + 1024;
static constexpr uint32_t PP_TMEM_COLS = 512;

struct AttnPpParams {
  CUtensorMap tma;  // qkv dims (3*dh, L, B), box (64, 128, 1)
  const float* bound_log2;
  __nv_bfloat16* y;
  float* lse;
  int B, H, L, dh;
  float scale_log2, scale;
};

__device__ __forceinline__ float pp_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one chunk of 32 score columns -> 32 probabilities (bf16 pairs in pk[16]); EMU of every 4 pairs use the FMA-pipe exponential
template <int EMU>
__device__ __forceinline__ void pp_chunk(const uint32_t (&r)[32], float c, float neg_mc, uint32_t (&pk)[16], float2& s01,
                                         float2& s23) {
  const float2 c2 = make_float2(c, c), n2 = make_float2(neg_mc, neg_mc);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float2 a = ffma2(make_float2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1])), c2, n2);
    float2 e;
    if ((p & 3) < EMU)
      e = ex2_poly2(a);
    else
      e = make_float2(pp_ex2(a.x), pp_ex2(a.y));
    if (p & 1)
      s23 = fadd2(s23, e);
    else
      s01 = fadd2(s01, e);
    pk[p] = pack_bf16(e.x, e.y);
  }
}
// same with the columns >= valid masked to zero (last kv tile of a ragged sequence)
__device__ __forceinline__ void pp_chunk_masked(const uint32_t (&r)[32], float c, float neg_mc, int valid, uint32_t (&pk)[16],
                                                float2& s01) {
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float e0 = (2 * p < valid) ? pp_ex2(fmaf(__uint_as_float(r[2 * p]), c, neg_mc)) : 0.f;
    const float e1 = (2 * p + 1 < valid) ? pp_ex2(fmaf(__uint_as_float(r[2 * p + 1]), c, neg_mc)) : 0.f;
    s01 = fadd2(s01, make_float2(e0, e1));
    pk[p] = pack_bf16(e0, e1);
  }
}

template <int EMU>
__global__ void __launch_bounds__(PP_THREADS, 1) attn_fwd_pp_kernel(const __grid_constant__ AttnPpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + PP_SMEM_TILES + 256 > dyn) __trap();
  }
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sK = sQ + 2 * PP_TILE;           // PP_STAGES tiles
  uint8_t* sV = sK + PP_STAGES * PP_TILE;   // PP_STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + PP_STAGES * PP_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2] S_g(j) complete
  uint64_t* p_full = bars + 15;   // [2] P_g(j) written by the 4 warps of group g
  uint64_t* o_ready = bars + 17;  // [2] O_g += P_g(j) V_j complete (online mode waits on it before rescaling O)
  uint64_t* acc_done = bars + 19; // [2] last PV of group g complete
  uint64_t* qt_ready = bars + 21; // Q tiles copied into TMEM by the 8 softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

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
    for (int i = 0; i < PP_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_ready[g], 1);
      mbar_init(&acc_done[g], 1);
    }
    mbar_init(qt_ready, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, PP_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * PP_TILE);
      tma_load_3d(sQ, &p.tma, q_full, h * 64, q0, b);
      tma_load_3d(sQ + PP_TILE, &p.tma, q_full, h * 64, q0 + 128, b);  // rows past L are zero-filled by TMA
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], PP_TILE);
        tma_load_3d(sK + st * PP_TILE, &p.tma, &k_full[st], p.dh + h * 64, j * 128, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], PP_TILE);
        tma_load_3d(sV + st * PP_TILE, &p.tma, &v_full[st], 2 * p.dh + h * 64, j * 128, b);
        if (++st == PP_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== UMMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(FMT_BF16, 0, 0, 128, 128);  // S = Q K^T : A in TMEM, B K-major, N = 128 kv
      const uint32_t idesc_o = make_idesc(FMT_BF16, 0, 1, 128, 64);   // O += P V : A in TMEM, B MN-major, N = 64 d
      auto issue_s = [&](int g, int j) {  // S_g(j) = Q_g K_j^T (K_j already waited for)
        const uint32_t aK = smem_u32(sK + (j % PP_STAGES) * PP_TILE);
        const uint32_t tS = tmem_base + g * 128, tQ = tmem_base + 384 + g * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tS, tQ + k * 8, make_smem_desc(aK + k * 32, 0, 1024), idesc_s, k > 0);
        umma_commit(&s_full[g]);
      };
      mbar_wait(qt_ready, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % PP_STAGES, nst = (j + 1) % PP_STAGES;
        mbar_wait(&v_full[st], (j / PP_STAGES) & 1);
        const uint32_t aV = smem_u32(sV + st * PP_TILE);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g], j & 1);
          tc_fence_after();
          const uint32_t tP = tmem_base + g * 128, tO = tmem_base + 256 + g * 64;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // contraction over the 128 kv rows, 16 per instruction
            umma_f16_ts(tO, tP + k * 8, make_smem_desc(aV + k * 16 * 128, 0, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&o_ready[g]);
          if (j + 1 == n_kv) umma_commit(&acc_done[g]);
          if (j + 1 < n_kv) {
            if (g == 0) {
              mbar_wait(&k_full[nst], ((j + 1) / PP_STAGES) & 1);
              tc_fence_after();
            }
            issue_s(g, j + 1);  // overwrites P_g(j): issued after the MMA that reads it
            if (g == 1) umma_commit(&k_empty[nst]);
          }
        }
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    // ================================================================== softmax group g (thread = q row of tile g)
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + g * 64 + lane_off;
    {  // one-time copy of this thread's Q row (128 B, SW128 smem) into TMEM
      mbar_wait(q_full, 0);
      const uint32_t base = smem_u32(sQ + g * PP_TILE) + row * 128;
      uint32_t rq[32];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rq[4 * u]), "=r"(rq[4 * u + 1]), "=r"(rq[4 * u + 2]), "=r"(rq[4 * u + 3])
                     : "r"(base + ((u ^ (row & 7)) << 4)));
      __syncwarp();
      tmem_st32(tmem_base + 384 + g * 32 + lane_off, rq);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qt_ready);
    }
    const float c = p.scale_log2;
    float bound = INFINITY;
    if (p.bound_log2 != nullptr) bound = __ldg(p.bound_log2);
    const bool fixed = bound < 3.0e38f;
    float m = fixed ? bound / c : -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.L - j * 128;
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      float alpha = 1.0f;
      if (!fixed) {  // online softmax: first pass over S for the row maximum, then rescale O if it moved
        float mx = -INFINITY;
#pragma unroll 1
        for (int cch = 0; cch < 4; ++cch) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tS + cch * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cch * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        const float m_new = fmaxf(m, mx);
        alpha = pp_ex2((m - m_new) * c);
        m = m_new;
        if (j > 0) {
          mbar_wait(&o_ready[g], (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
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
      }
      const float neg_mc = -m * c;
      float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
      uint32_t rb[2][32];
      __syncwarp();
      tmem_ld32(tS, rb[0]);
#pragma unroll
      for (int cch = 0; cch < 4; ++cch) {
        tmem_wait_ld();
        if (cch < 3) {
          __syncwarp();
          tmem_ld32(tS + (cch + 1) * 32, rb[(cch + 1) & 1]);  // next chunk streams out of TMEM behind this chunk's math
        }
        uint32_t pk[16];
        if (valid >= 128)
          pp_chunk<EMU>(rb[cch & 1], c, neg_mc, pk, s01, s23);
        else
          pp_chunk_masked(rb[cch & 1], c, neg_mc, valid - cch * 32, pk, s01);
        tmem_st16(tS + cch * 16, pk);  // P over S columns this thread has already consumed
      }
      tmem_wait_st();
      l = l * alpha + ((s01.x + s01.y) + (s23.x + s23.y));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
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
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, PP_TMEM_COLS);
  }
}

template <int EMU>
static int launch_pp_t(const AttnPpParams& p, long long grid, cudaStream_t stream) {
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_BYTES));
  }
  attn_fwd_pp_kernel<EMU><<<(unsigned)grid, PP_THREADS, PP_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  return 0;
}

// emu: share of the exponentials on the FMA pipe, in quarters (0, 1, 2); < 0 = the default (OSD_PP_EMU or 1)
int launch_attn_fwd_pp(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H, int emu,
                       cudaStream_t stream) {
  OSD_CHECK(qkv && y && B > 0 && L > 0 && H > 0, "attn_fwd_pp: bad arguments");
  AttnPpParams p;
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
    return (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }();
  if (emu < 0) emu = emu_default;
  const long long grid = (long long)ceil_div(L, 256) * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_fwd_pp: grid too large");
  if (emu == 0) return launch_pp_t<0>(p, grid, stream);
  if (emu == 2) return launch_pp_t<2>(p, grid, stream);
  return launch_pp_t<1>(p, grid, stream);
}

}  // namespace osd
