// fp32-grade flash attention forward: every bf16 tensor-core product is replaced by the split
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo      (a = a_hi + a_lo, both bf16: ~16 mantissa bits)
// so that S = Q K^T and O = P V carry fp32-like accuracy while still running on tcgen05 (which has no fp32 MMA and
// whose kind::tf32 truncates its operands).  Used by the precision='fp32' sampling path (BASELINE config 3).
// Which correction terms run is a launch parameter (qk_mask / pv_mask: bit 1 = lo*hi, bit 2 = hi*lo): the attention
// kernels are ENERGY bound on the 1 kW part, so every dropped term is time; the error each term buys
// is measured by tools/x3_terms_sweep.py and the shipped choice is the cheapest one that keeps the 64-step sampler a
// decade inside the 1e-3 tolerance.  When P_lo is not multiplied, the row sum is taken over the ROUNDED P_hi, so that
// the normalisation sees the same probabilities as the MMA (a row dominated by one key stays exact).
// Same structure as attn_fwd_kernel<64, true>: 128 q rows per CTA, 64-row kv tiles, P kept in TMEM.
//   qkv : bf16 [B*L, 2*3*dh] = (hi block | lo block), each block (q | k | v)
//   y   : bf16 [B*L, 2*dh]   = (hi | lo)
//   TMEM: S fp32 cols [0,64) -- P_hi (bf16) is written back over cols [0,32), P_lo over [32,64) -- | O [64,128)
#include "kernels.cuh"
#include "ptx.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace osd {

// shipped term selection: all three terms for both products until tools/x3_terms_sweep.py says otherwise
#ifndef X3_DEFAULT_QK
#define X3_DEFAULT_QK 7
#endif
#ifndef X3_DEFAULT_PV
#define X3_DEFAULT_PV 7
#endif
static constexpr int X3_THREADS = 192;
static constexpr int X3_T128 = 128 * 128;
static constexpr int X3_T64 = 64 * 128;
static constexpr int X3_SMEM_TILES = 2 * X3_T128 + 8 * X3_T64;  // Qh Ql | (Kh Kl) x2 | (Vh Vl) x2
static constexpr int X3_SMEM_BYTES = X3_SMEM_TILES + 256;
static constexpr uint32_t X3_TMEM_COLS = 128;

struct AttnX3Params {
  CUtensorMap tma_q;   // dims (6*dh, L, B), box (64, 128, 1)
  CUtensorMap tma_kv;  // same tensor, box (64, 64, 1)
  const float* bound_log2;
  __nv_bfloat16* y;  // [B*L, 2*dh]
  float* lse;
  int B, H, L, dh;
  float scale_log2, scale;
  int qk_mask, pv_mask;  // bit 1: (Q_lo K_hi | P_lo V_hi), bit 2: (Q_hi K_lo | P_hi V_lo); the hi*hi term always runs
};

__device__ __forceinline__ float x3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(X3_THREADS, 2) attn_fwd_x3_kernel(const __grid_constant__ AttnX3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + X3_SMEM_TILES + 128 > dyn) __trap();
  }
  uint8_t* sQ = smem;               // Qh | Ql
  uint8_t* sK = sQ + 2 * X3_T128;   // stage s: Kh | Kl
  uint8_t* sV = sK + 4 * X3_T64;    // stage s: Vh | Vl
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 4 * X3_T64);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_ready = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (p.L + 127) / 128;
  const int qt = blockIdx.x % n_qt;
  const int bh = blockIdx.x / n_qt;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * 128;
  const int n_kv = (p.L + 63) / 64;
  const int lo_col = 3 * p.dh;  // start of the lo block

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, X3_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, ((p.qk_mask & 2) ? 2 : 1) * X3_T128);
      tma_load_3d(sQ, &p.tma_q, q_full, h * 64, q0, b);
      if (p.qk_mask & 2) tma_load_3d(sQ + X3_T128, &p.tma_q, q_full, lo_col + h * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        const bool klo = (p.qk_mask & 4) != 0, vlo = (p.pv_mask & 4) != 0;  // the lo tiles travel only when a term reads them
        mbar_expect_tx(&k_full[st], (klo ? 2 : 1) * X3_T64);
        tma_load_3d(sK + (2 * st) * X3_T64, &p.tma_kv, &k_full[st], p.dh + h * 64, j * 64, b);
        if (klo) tma_load_3d(sK + (2 * st + 1) * X3_T64, &p.tma_kv, &k_full[st], lo_col + p.dh + h * 64, j * 64, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], (vlo ? 2 : 1) * X3_T64);
        tma_load_3d(sV + (2 * st) * X3_T64, &p.tma_kv, &v_full[st], 2 * p.dh + h * 64, j * 64, b);
        if (vlo) tma_load_3d(sV + (2 * st + 1) * X3_T64, &p.tma_kv, &v_full[st], lo_col + 2 * p.dh + h * 64, j * 64, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(FMT_BF16, 0, 0, 128, 64);
      const uint32_t idesc_o = make_idesc(FMT_BF16, 0, 1, 128, 64);
      const uint32_t aQh = smem_u32(sQ), aQl = aQh + X3_T128;
      const uint32_t tS = tmem_base, tO = tmem_base + 64;
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aKh = smem_u32(sK + (2 * st) * X3_T64), aKl = aKh + X3_T64;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // Q_hi K_hi^T
          umma_f16_ss(tS, make_smem_desc(aQh + k * 32, 0, 1024), make_smem_desc(aKh + k * 32, 0, 1024), idesc_s, k > 0);
        if (p.qk_mask & 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // Q_lo K_hi^T
            umma_f16_ss(tS, make_smem_desc(aQl + k * 32, 0, 1024), make_smem_desc(aKh + k * 32, 0, 1024), idesc_s, 1u);
        }
        if (p.qk_mask & 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // Q_hi K_lo^T
            umma_f16_ss(tS, make_smem_desc(aQh + k * 32, 0, 1024), make_smem_desc(aKl + k * 32, 0, 1024), idesc_s, 1u);
        }
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aVh = smem_u32(sV + (2 * st) * X3_T64), aVl = aVh + X3_T64;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // P_hi V_hi
          umma_f16_ts(tO, tS + k * 8, make_smem_desc(aVh + k * 16 * 128, 0, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        if (p.pv_mask & 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // P_lo V_hi
            umma_f16_ts(tO, tS + 32 + k * 8, make_smem_desc(aVh + k * 16 * 128, 0, 1024), idesc_o, 1u);
        }
        if (p.pv_mask & 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // P_hi V_lo
            umma_f16_ts(tO, tS + k * 8, make_smem_desc(aVl + k * 16 * 128, 0, 1024), idesc_o, 1u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(o_ready);
        if (j + 1 < n_kv) issue_s(j + 1);  // in-order tensor pipe: overwrites P only after the MMAs above
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off, tO = tmem_base + 64 + lane_off;
    const float c = p.scale_log2;
    float bound = INFINITY;
    if (p.bound_log2 != nullptr) bound = __ldg(p.bound_log2);
    const bool fixed = bound < 3.0e38f;
    float m = fixed ? bound / c : -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int valid = p.L - j * 64;
      uint32_t r0[32], r1[32];
      __syncwarp();
      tmem_ld32(tS, r0);
      tmem_ld32(tS + 32, r1);
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
        alpha = x3_ex2((m - m_new) * c);
      }
      const float neg_mc = -m_new * c;
      if (j > 0) {
        mbar_wait(o_ready, (j - 1) & 1);
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
      float sum = 0.f;
      const bool plo = (p.pv_mask & 2) != 0;
      uint32_t ph[32], pl[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float a0 = (i < valid) ? x3_ex2(fmaf(__uint_as_float(r0[i]), c, neg_mc)) : 0.f;
        float a1 = (i + 1 < valid) ? x3_ex2(fmaf(__uint_as_float(r0[i + 1]), c, neg_mc)) : 0.f;
        float b0 = (32 + i < valid) ? x3_ex2(fmaf(__uint_as_float(r1[i]), c, neg_mc)) : 0.f;
        float b1 = (33 + i < valid) ? x3_ex2(fmaf(__uint_as_float(r1[i + 1]), c, neg_mc)) : 0.f;
        const uint32_t ha = pack_bf16(a0, a1), hb = pack_bf16(b0, b1);
        const __nv_bfloat162 fa = *reinterpret_cast<const __nv_bfloat162*>(&ha);
        const __nv_bfloat162 fb = *reinterpret_cast<const __nv_bfloat162*>(&hb);
        if (plo)
          sum += (a0 + a1) + (b0 + b1);
        else  // only P_hi is multiplied: normalise by what the MMA sees
          sum += (__low2float(fa) + __high2float(fa)) + (__low2float(fb) + __high2float(fb));
        ph[i >> 1] = ha;
        ph[16 + (i >> 1)] = hb;
        pl[i >> 1] = pack_bf16(a0 - __low2float(fa), a1 - __high2float(fa));
        pl[16 + (i >> 1)] = pack_bf16(b0 - __low2float(fb), b1 - __high2float(fb));
      }
      __syncwarp();
      tmem_st32(tS, ph);                // P_hi: 64 bf16 = 32 columns
      if (plo) tmem_st32(tS + 32, pl);  // P_lo
      tmem_wait_st();
      l = l * alpha + sum;
      m = m_new;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_ready, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int q = q0 + row;
    const bool ok = q < p.L;
#pragma unroll 1
    for (int cch = 0; cch < 2; ++cch) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tO + cch * 32, r);
      tmem_wait_ld();
      if (ok) {
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
            const __nv_bfloat162 hf = *reinterpret_cast<const __nv_bfloat162*>(&hw[e]);
            lw[e] = pack_bf16(v0 - __low2float(hf), v1 - __high2float(hf));
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
    tmem_dealloc(tmem_base, X3_TMEM_COLS);
  }
}

// OSD_X3_TERMS="<qk_mask>,<pv_mask>" overrides the shipped term selection (tools/x3_terms_sweep.py)
static void x3_terms(int* qk, int* pv) {
  static const int packed = [] {
    int a = X3_DEFAULT_QK, b = X3_DEFAULT_PV;
    const char* e = getenv("OSD_X3_TERMS");
    if (e != nullptr) sscanf(e, "%d,%d", &a, &b);
    return ((a | 1) & 7) | (((b | 1) & 7) << 8);
  }();
  *qk = packed & 7;
  *pv = (packed >> 8) & 7;
}

int launch_attn_fwd_x3(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                       cudaStream_t stream) {
  OSD_CHECK(qkv && y && B > 0 && L > 0 && H > 0, "attn_fwd_x3: bad arguments");
  AttnX3Params p;
  const int dh = H * 64;
  uint64_t dims[3] = {(uint64_t)6 * dh, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)6 * dh * 2, (uint64_t)L * 6 * dh * 2};
  uint32_t box_q[3] = {64, 128, 1}, box_kv[3] = {64, 64, 1};
  OSD_TRY(make_tmap(&p.tma_q, qkv, 2, 3, dims, strides, box_q));
  OSD_TRY(make_tmap(&p.tma_kv, qkv, 2, 3, dims, strides, box_kv));
  p.bound_log2 = bound_log2;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.lse = lse;
  p.B = B; p.H = H; p.L = L; p.dh = dh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  x3_terms(&p.qk_mask, &p.pv_mask);
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(attn_fwd_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM_BYTES));
  }
  const long long grid = (long long)ceil_div(L, 128) * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_fwd_x3: grid too large");
  attn_fwd_x3_kernel<<<(unsigned)grid, X3_THREADS, X3_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
