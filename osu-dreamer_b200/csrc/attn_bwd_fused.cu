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
//   warps    : 0 TMA producer | 1 TMEM allocator + UMMA issuer | 2-9 softmax (2 column groups x 4 lane quadrants,
//              thread = kv row); the same warps drain the finished dQ tiles (TMEM -> swizzled smem -> TMA
//              reduce-add; warp = 32 q rows x 32 d columns).  10 warps keep the register file at <= 3 warps per SM
//              sub-partition (200 registers per thread available).  The two column groups have separate barriers
//              and run half a tile period apart (see the UMMA issue order), so one group's exponentials (SFU) overlap
//              the other group's dS phase.
//   TMEM     : S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ_i [384,448) | K bf16 [448,480) |
//              V bf16 [480,512).  K and V are the A operands of S^T = K Q_i^T and dP^T = V dO_i^T straight from TMEM;
//              the bf16 P^T / dS^T tiles are written back over the consumed S^T / dP^T columns (column group g at
//              +64g) and are the TMEM A operands of dV += P^T dO_i and dK += dS^T Q_i.
//   smem     : K | V (staging for the TMEM copy, then dQ staging; K also B operand of dQ) | Q_i x3 | dO_i x3 |
//              dS^T x2 (bf16, MN-major A operand of dQ_i = dS_i K) | -lse2 / -D x3 (bulk-copied with Q_i) | barriers
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

static constexpr int FB_THREADS = 320;
static constexpr int FT = 128 * 128;  // bytes of a [128 x 64] bf16 tile
static constexpr uint32_t FB_TMEM_COLS = 512;
static constexpr int FB_QST = 3;                      // Q / dO / statistics stages
static constexpr int FB_STAT_BYTES = FB_QST * 1024;   // per stage: [-lse2 128 | -D 128] fp32
// smem map (bytes): K 16K | V 16K + 16K (V is only staging for the TMEM copy; afterwards these 32K are the dQ
// staging tiles, 8 warps x [32 rows x 128 B]) | Q x3 | dO x3 | dS^T x2 (each 2 atoms of [128 kv x 64 q] bf16) | stats
static constexpr int FB_OFF_V = FT;
static constexpr int FB_OFF_Q = 3 * FT;
static constexpr int FB_OFF_DO = FB_OFF_Q + FB_QST * FT;
static constexpr int FB_OFF_DS = FB_OFF_DO + FB_QST * FT;
static constexpr int FB_OFF_STAT = FB_OFF_DS + 4 * FT;
static constexpr int FB_SMEM_TILES = FB_OFF_STAT + FB_STAT_BYTES;
static constexpr int FB_SMEM_BYTES = FB_SMEM_TILES + 256 + 1024;

__device__ __forceinline__ float fb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnBwdFusedParams {
  CUtensorMap tma_qkv;  // qkv    dims (3*dh, L, B), box (64, 128, 1)
  CUtensorMap tma_dy;   // dy     dims (dh, L, B),   box (64, 128, 1)
  CUtensorMap tma_dq;   // dq_acc dims (dh, L, B) fp32, box (32, 32, 1)
  const float* lse2n;   // [B*H, Lp]  -lse * log2(e), zero past L   (Lp = L rounded up to 128)
  const float* dneg;    // [B*H, Lp]  -D = -rowsum(dO o O), zero past L
  __nv_bfloat16* dqkv;  // [B*L, 3*dh]: this kernel writes the dk and dv column blocks
  int B, H, L, Lp, dh;
  float scale, scale_log2;
};

__global__ void __launch_bounds__(FB_THREADS, 1) attn_bwd_fused_kernel(const __grid_constant__ AttnBwdFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  {
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (pad + FB_SMEM_TILES + 256 > dyn) __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = smem + FB_OFF_V;
  uint8_t* sStg = sV;  // aliases V (+16K): V is dead once it has been copied into TMEM
  uint8_t* sQ = smem + FB_OFF_Q;
  uint8_t* sDO = smem + FB_OFF_DO;
  uint8_t* sDS = smem + FB_OFF_DS;  // [2 buffers][2 q-chunks][128 kv rows][64 q] bf16
  float* sStat = reinterpret_cast<float*>(smem + FB_OFF_STAT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FB_SMEM_TILES);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // [3]
  uint64_t* q_empty = bars + 4;   // [3]
  uint64_t* s_full = bars + 7;    // [2 groups]  S^T_i (column group g) complete
  uint64_t* dp_full = bars + 9;   // [2]         dP^T_i (group g) complete
  uint64_t* pt_full = bars + 11;  // [2]         P^T_i (group g) written to TMEM
  uint64_t* ds_full = bars + 13;  // [2]         dS^T_i (group g) written to TMEM + smem
  uint64_t* dq_full = bars + 15;  // dQ_i complete
  uint64_t* dq_empty = bars + 16; // dQ_i drained out of TMEM
  uint64_t* kvt_ready = bars + 17;  // K / V copied into TMEM
  uint64_t* acc_done = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_t = (p.L + 127) / 128;  // kv tiles == q tiles
  const int kt = blockIdx.x % n_t;
  const int bh = blockIdx.x / n_t;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kt * 128;
  const int n_q = n_t;
  const int i0 = kt;  // q-loop rotation: iteration i works on q tile (i0 + i) % n_q

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    tma_prefetch_desc(&p.tma_dy);
    tma_prefetch_desc(&p.tma_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < FB_QST; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&dp_full[g], 1);
      mbar_init(&pt_full[g], 4);
      mbar_init(&ds_full[g], 4);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    mbar_init(kvt_ready, 8);
    mbar_init(acc_done, 1);
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
      int qi = i0, st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(&q_empty[st], ph ^ 1);
        mbar_expect_tx(&q_full[st], 2 * FT + 1024);
        tma_load_3d(sQ + st * FT, &p.tma_qkv, &q_full[st], h * 64, qi * 128, b);
        tma_load_3d(sDO + st * FT, &p.tma_dy, &q_full[st], h * 64, qi * 128, b);
        const size_t so = ((size_t)b * p.H + h) * p.Lp + (size_t)qi * 128;
        bulk_load_1d(sStat + st * 256, p.lse2n + so, 512, &q_full[st]);
        bulk_load_1d(sStat + st * 256 + 128, p.dneg + so, 512, &q_full[st]);
        if (++qi == n_q) qi = 0;
        if (++st == FB_QST) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== UMMA issuer
    // Column group g (q columns [64g, 64g+64) of the tile) has its own S^T / dP^T halves and barriers.  Within a
    // group the in-order tensor pipe makes the TMEM aliasing (P^T over S^T, dS^T over dP^T) safe: dV(i,g) is issued
    // before S(i+1,g), dK(i,g) before dP(i+1,g).
    if (elect_one()) {
      const uint32_t id_s = make_idesc(FMT_BF16, 0, 0, 128, 64);  // S^T, dP^T halves: A in TMEM, B K-major, N = 64 q
      const uint32_t id_o = make_idesc(FMT_BF16, 0, 1, 128, 64);  // dV, dK: A in TMEM, B MN-major, N = 64 d
      const uint32_t id_q = make_idesc(FMT_BF16, 1, 1, 128, 64);  // dQ: A = dS MN-major smem, B = K MN-major
      const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320;
      const uint32_t tDQ = tmem_base + 384, tK = tmem_base + 448, tV = tmem_base + 480;
      const uint32_t aK = smem_u32(sK), aDS = smem_u32(sDS), aQ0 = smem_u32(sQ), aDO0 = smem_u32(sDO);
      uint32_t dv_acc = 0, dk_acc = 0;
      // tile i lives in stage i % 3
      auto issue_s = [&](int g, int st) {  // S^T(.,g) = K Q^T over q rows [64g, +64)
        const uint32_t aQ = aQ0 + st * FT + g * 8192;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tS + g * 64, tK + k * 8, make_smem_desc(aQ + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(&s_full[g]);
      };
      auto issue_dp = [&](int g, int st) {  // dP^T(.,g) = V dO^T
        const uint32_t aDO = aDO0 + st * FT + g * 8192;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tDP + g * 64, tV + k * 8, make_smem_desc(aDO + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(&dp_full[g]);
      };
      auto issue_dv = [&](int g, int st) {  // dV += P^T(.,g) dO[64g.., :]
        const uint32_t aDO = aDO0 + st * FT + g * 8192;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_f16_ts(tDV, tS + g * 64 + k * 8, make_smem_desc(aDO + k * 16 * 128, 0, 1024), id_o, dv_acc);
          dv_acc = 1;
        }
      };
      auto issue_dk = [&](int g, int st) {  // dK += dS^T(.,g) Q[64g.., :]
        const uint32_t aQ = aQ0 + st * FT + g * 8192;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_f16_ts(tDK, tDP + g * 64 + k * 8, make_smem_desc(aQ + k * 16 * 128, 0, 1024), id_o, dk_acc);
          dk_acc = 1;
        }
      };
      // Event-driven issue: each column group's chain  [P^T_i] dV(i,g) S(i+1,g)  ->  [dS^T_i] dK(i,g) dP(i+1,g)  advances
      // as soon as its own barrier completes (polled, never blocking the other group); dQ_j is issued once both
      // groups have produced dS^T_j and dQ_{j-1} has left TMEM.  Group 1's first tile is held back until group 0 has
      // finished its first exponential phase, which puts the two groups about half a tile period apart.
      mbar_wait(kvt_ready, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_dp(0, 0);
      int ti[2] = {0, 0};        // tile whose event group g is waiting for
      int stg_of[2] = {0, 0};    // stage of that tile
      uint32_t qph[2] = {0, 0};  // q_full parity of that stage
      bool want_ds[2] = {false, false};
      bool s_pend[2] = {false, false};  // S(i+1,g) waiting for its Q stage (never blocks the other group's chain)
      int s_st[2] = {0, 0};
      uint32_t s_ph[2] = {0, 0};
      bool started1 = false;
      int dk_done[2] = {0, 0};  // tiles whose dK (last reader of the Q / dO stage) has been issued
      int released = 0, rel_st = 0, dq_next = 0;
      while (ti[0] < n_q || ti[1] < n_q || dq_next < n_q) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (s_pend[g] && mbar_try_wait(&q_full[s_st[g]], s_ph[g])) {
            tc_fence_after();
            issue_s(g, s_st[g]);
            s_pend[g] = false;
          }
          if (ti[g] >= n_q || (g == 1 && !started1)) continue;
          const int i = ti[g], st = stg_of[g];
          if (!want_ds[g]) {
            if (!mbar_try_wait(&pt_full[g], i & 1)) continue;
            tc_fence_after();
            issue_dv(g, st);
            if (i + 1 < n_q) {  // S(i+1,g) follows as soon as Q_{i+1} has landed (polled below)
              s_pend[g] = true;
              s_st[g] = st + 1;
              s_ph[g] = qph[g];
              if (s_st[g] == FB_QST) {
                s_st[g] = 0;
                s_ph[g] ^= 1;
              }
            }
            want_ds[g] = true;
            if (g == 0 && !started1) {
              started1 = true;
              issue_s(1, 0);
              issue_dp(1, 0);
            }
          } else {
            if (!mbar_try_wait(&ds_full[g], i & 1)) continue;
            tc_fence_after();
            issue_dk(g, st);
            int sn = st + 1;
            if (sn == FB_QST) {
              sn = 0;
              qph[g] ^= 1;
            }
            if (i + 1 < n_q) issue_dp(g, sn);
            dk_done[g] = i + 1;
            stg_of[g] = sn;
            ti[g] = i + 1;
            want_ds[g] = false;
          }
        }
        const int both = dk_done[0] < dk_done[1] ? dk_done[0] : dk_done[1];
        if (released < both) {  // every MMA that reads stage rel_st has been issued
          umma_commit(&q_empty[rel_st]);
          ++released;
          if (++rel_st == FB_QST) rel_st = 0;
        }
        if (dq_next < both && (dq_next == 0 || mbar_try_wait(dq_empty, (dq_next - 1) & 1))) {
          tc_fence_after();
          const uint32_t a = aDS + (dq_next & 1) * 2 * FT;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dQ_j = dS_j K : contraction over the 128 kv rows, 16 per instruction
            umma_f16_ss(tDQ, make_smem_desc(a + k * 16 * 128, FT, 1024), make_smem_desc(aK + k * 16 * 128, 0, 1024), id_q,
                        k > 0);
          umma_commit(dq_full);
          ++dq_next;
        }
      }
      umma_commit(acc_done);
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
    uint64_t* my_s_full = &s_full[grp];
    uint64_t* my_dp_full = &dp_full[grp];
    uint64_t* my_pt_full = &pt_full[grp];
    uint64_t* my_ds_full = &ds_full[grp];
    // dQ drain: this warp moves rows [32 quad, +32) x columns [32 grp, +32) of the finished dQ tile j (q tile qj):
    // part 1 = TMEM -> registers -> staging smem (the proxy fence is shared with the dS^T smem writes of phase 2),
    // part 2 (after that fence) = one lane issues the TMA reduce-add
    const uint32_t tDQ = tmem_base + 384 + lane_off + grp * 32;
    uint8_t* stg = sStg + (warp - 2) * 4096;
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
      if (lane == 0) {
        tma_reduce_add_3d(&p.tma_dq, stg, h * 64 + grp * 32, qj * 128 + quad * 32, b);
        tma_store_commit();
      }
    };
    // Drain schedule: dQ_{i-2} is drained between the two phases of tile i (it completed about a tile period ago, so
    // the wait never blocks; it also guarantees that dS^T smem buffer i&1, last read by dQ_{i-2}, is free).
    int qt = i0, qprev = i0, qprev2 = i0, stq = 0;
    for (int i = 0; i < n_q; ++i) {
      // [-lse2 128 | -D 128] of this q tile, bulk-copied on the same barrier as Q_i / dO_i (complete before S^T_i)
      const float* st = sStat + stq * 256 + grp * 64;
      // ---- phase 1: P^T = exp2(S^T * c - lse2[q])
      mbar_wait(my_s_full, i & 1);
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
          const float4 lv = l4[k4];
          const float2 a = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 0]), __uint_as_float(rs[k4 * 4 + 1])), c2,
                                 make_float2(lv.x, lv.y));
          const float2 bb = ffma2(make_float2(__uint_as_float(rs[k4 * 4 + 2]), __uint_as_float(rs[k4 * 4 + 3])), c2,
                                  make_float2(lv.z, lv.w));
          pt[cch * 32 + k4 * 4 + 0] = fb_ex2(a.x);
          pt[cch * 32 + k4 * 4 + 1] = fb_ex2(a.y);
          pt[cch * 32 + k4 * 4 + 2] = fb_ex2(bb.x);
          pt[cch * 32 + k4 * 4 + 3] = fb_ex2(bb.y);
        }
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) pk[k] = pack_bf16(pt[cch * 32 + 2 * k], pt[cch * 32 + 2 * k + 1]);
        tmem_st16(tS + cch * 16, pk);  // over S^T columns this thread has already consumed
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(my_pt_full);
      if (i > 1) drain_load(i - 2);
      // ---- phase 2: dS^T = P^T o (dP^T - D[q])   (1/sqrt(d) is applied to the dK / dQ results on the way out)
      mbar_wait(my_dp_full, i & 1);
      tc_fence_after();
      const uint32_t ds_dst = ds_row + (i & 1) * 2 * FT;  // buffer i&1: dQ_{i-2} (its last reader) was drained by this warp
#pragma unroll
      for (int cch = 0; cch < 2; ++cch) {
        uint32_t rp[32];
        __syncwarp();
        tmem_ld32(tDP + cch * 32, rp);
        tmem_wait_ld();
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
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_dst + (((cch * 4 + u4) ^ sw) << 4)),
                       "r"(pk[4 * u4]), "r"(pk[4 * u4 + 1]), "r"(pk[4 * u4 + 2]), "r"(pk[4 * u4 + 3])
                       : "memory");
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(my_ds_full);
      if (i > 1) drain_issue(qprev2);
      qprev2 = qprev;
      qprev = qt;
      if (++qt == n_q) qt = 0;
      if (++stq == FB_QST) stq = 0;
    }
    if (n_q > 1) {
      drain_load(n_q - 2);
      fence_proxy_async_smem();
      __syncwarp();
      drain_issue(qprev2);
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

// stats[0] = -lse * log2(e), stats[1] = -rowsum(dO o O), both [B*H, Lp] with zeros for Lp > l >= L
// (warp per padded token: lane covers 32 contiguous columns of the 1024 = half a head)
__global__ void attn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                                      const float* __restrict__ lse, float* __restrict__ lse2n, float* __restrict__ dneg,
                                      int B, int H, int L, int Lp) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= B * Lp) return;
  const int b = t / Lp, l = t % Lp;
  float acc = 0.f;
  if (l < L) {
    const size_t tok = (size_t)b * L + l;
    const uint4* a = reinterpret_cast<const uint4*>(y + tok * (H * 64) + lane * 32);
    const uint4* g = reinterpret_cast<const uint4*>(dy + tok * (H * 64) + lane * 32);
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
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if ((lane & 1) == 0) {
    const int hh = lane >> 1;
    const size_t o = ((size_t)b * H + hh) * Lp + l;
    dneg[o] = -acc;
    lse2n[o] = (l < L) ? -lse[((size_t)b * H + hh) * L + l] * 1.4426950408889634f : 0.f;
  }
}

// dq_acc fp32 [T, dh] (unscaled) -> bf16 dq column block of dqkv [T, 3*dh], x 1/sqrt(d)
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, size_t n8,
                                           int dh, float scale) {
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

size_t attn_bwd_fused_stats_floats(int B, int L, int H) { return (size_t)2 * B * H * ((L + 127) / 128 * 128); }

int launch_attn_bwd_fused(const void* qkv, const void* y, const void* dy, const float* lse, float* stats, float* dq_acc,
                          void* dqkv, int B, int L, int H, cudaStream_t stream) {
  OSD_CHECK(qkv && y && dy && lse && stats && dq_acc && dqkv && B > 0 && L > 0 && H == 16, "attn_bwd_fused: bad arguments");
  const int dh = H * 64;
  const int Lp = (L + 127) / 128 * 128;
  float* lse2n = stats;
  float* dneg = stats + (size_t)B * H * Lp;
  attn_bwd_stats_kernel<<<ceil_div(B * Lp, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y),
                                                                static_cast<const __nv_bfloat16*>(dy), lse, lse2n, dneg, B,
                                                                H, L, Lp);
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
    OSD_TRY(make_tmap(&p.tma_dy, dy, 2, 3, dims, strides, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)dh, (uint64_t)L, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)dh * 4, (uint64_t)L * dh * 4};
    uint32_t box[3] = {32, 32, 1};
    OSD_TRY(make_tmap(&p.tma_dq, dq_acc, 4, 3, dims, strides, box));
  }
  p.lse2n = lse2n;
  p.dneg = dneg;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.B = B; p.H = H; p.L = L; p.Lp = Lp; p.dh = dh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    OSD_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES));
    attr_set = true;
  }
  const long long grid = (long long)ceil_div(L, 128) * H * B;
  OSD_CHECK(grid < (1ll << 31), "attn_bwd_fused: grid too large");
  attn_bwd_fused_kernel<<<(unsigned)grid, FB_THREADS, FB_SMEM_BYTES, stream>>>(p);
  OSD_LAUNCHED();
  const size_t n8 = (size_t)B * L * dh / 8;
  attn_bwd_dq_convert_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(dq_acc, p.dqkv, n8, dh, p.scale);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
