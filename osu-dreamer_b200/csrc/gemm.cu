// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + UMMA issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM -> registers -> fused epilogue -> global), lane quadrant = warp % 4
//                 (warps 2..9 = two per quadrant where the main loop is short, EW = 8)
// Pipelines: smem full/empty ring (TMA <-> UMMA) and a 2-deep TMEM accumulator ring (UMMA <-> epilogue),
// so the epilogue of work item i overlaps the main loop of item i+1.
// PAIR instantiation: the same roles in both CTAs of a 2-CTA cluster, one tcgen05.mma.cta_group::2 (M = 256) issued by the
// leader over both CTAs' shared memory and TMEM; see Cfg and the helpers under "CTA pairs" in ptx.cuh.
#include "gemm.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace osd {

static constexpr int BM = 128;
// threads = 64 (producer + issuer warps) + 32 per epilogue warp (EW = 4 or 8)
static constexpr int ROW_BYTES = 128;  // one swizzle row: 64 bf16 or 32 tf32

struct GemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  CUtensorMap tma_c;    // output, box (128 bytes, 32 rows), SW128
  CUtensorMap tma_raw;  // EPI_QKV: second output (pre-norm projections)
  int M, N, K;
  int a_major, b_major;
  int num_kb;        // total k-blocks
  int split3;        // 3x-bf16: num_kb = 3 * kb_seg, operands (hi | lo) along K
  int kb_seg;        // k-blocks of one K-long segment
  int c_split;       // bf16 output as (hi | lo)
  int kb_per_split;  // k-blocks per split
  int split_k;
  int tiles_m, tiles_n;
  int epi;
  void* C;
  long long ldc;
  int c_fp32;
  const float* bias;
  const float* qnorm_w;
  const float* knorm_w;
  const float* rope;
  int L;
  int dh;
  void* raw_out;
};

// EW = epilogue warps.  With 4 (one per SM sub-partition) the epilogue is a single latency-bound instruction stream per
// sub-partition; short-K GEMMs (K <= 1024: qkv, vg, the dgrads) are epilogue-bound that way, so they run with 8 (two
// warps per lane quadrant, alternating column chunks) and pay for the extra staging with one pipeline stage.
// PAIR: the work item is a [256 x BN] tile computed by the two CTAs of a cluster with cta_group::2 MMAs (ptx.cuh): each CTA
// stages its own 128 A rows and HALF of the B tile, so a stage is BN/2 rows smaller (more stages fit) and the tensor core's
// shared-memory operand reads per FLOP drop by a third (A 4 KB + B 4 KB instead of 4 + 8 per K = 16 step at BN = 256).
template <int BN, int ELEM, int EW, bool PAIR = false>
struct Cfg {
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int EPR = (ELEM == ELEM_BF16) ? 64 : 32;  // elements per 128-byte row
  static constexpr int BK = EPR;                             // k-extent of one stage (elements)
  static constexpr int UMMA_K = EPR / 4;                     // 32 bytes of K per instruction
  static constexpr int TM = PAIR ? 2 * BM : BM;              // rows of a work item
  static constexpr int BNL = PAIR ? BN / 2 : BN;             // B rows staged by one CTA
  static constexpr int A_BYTES = BM * ROW_BYTES;             // 16 KB
  static constexpr int B_BYTES = BNL * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = PAIR ? ((BN == 256) ? (EW == 8 ? 5 : 6) : (EW == 8 ? 6 : 8))
                                     : ((BN == 256) ? (EW == 8 ? 3 : 4) : (EW == 8 ? 5 : 6));
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulators (power of two: 256 or 512)
  static constexpr int STG_BYTES = EW * 2 * 4096;  // epilogue staging: EW warps x 2 buffers x [32 rows x 128 B]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------------------------------------------
template <int BN, int ELEM, bool QKV, int EW, bool PAIR>
__global__ void __launch_bounds__(64 + 32 * EW, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, ELEM, EW, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stg_base = smem + C::STAGES * C::STAGE_BYTES;  // 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + C::STG_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tfull_bar = bars + 2 * C::STAGES;
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader: issues the pair's MMAs, owns full / tempty barriers
  const int worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int n_workers = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    tma_prefetch_desc(&p.tma_c);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * EW : EW);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();  // (also in PAIR mode: compute-sanitizer's racecheck does not model barrier.cluster as a CTA barrier and
                    //  reports the TMEM allocator's shared-memory write against the read below)
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_items = p.tiles_m * p.tiles_n * p.split_k;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = worker; item < total_items; item += n_workers) {
        const int split = item % p.split_k;
        const int tile = item / p.split_k;
        const int tn = tile % p.tiles_n, tm = tile / p.tiles_n;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int row0 = tm * C::TM + (int)rank * BM;         // this CTA's A rows
        const int col0 = tn * BN + (int)rank * C::BNL;        // this CTA's B rows (all of them, or its half)
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          // PAIR: both CTAs' bytes complete on the leader's barrier
          const uint32_t fbar = PAIR ? mapa_shared(smem_u32(&full_bar[stage]), 0) : 0u;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], PAIR ? 2 * C::STAGE_BYTES : C::STAGE_BYTES);
          auto load = [&](uint8_t* dst, const CUtensorMap* map, int c0, int c1) {
            if constexpr (PAIR)
              tma_load_2d_2sm(dst, map, fbar, c0, c1);
            else
              tma_load_2d(dst, map, &full_bar[stage], c0, c1);
          };
          int k0 = kb * C::BK, k0b = k0;
          if (p.split3) {  // segments: (A_hi, B_hi), (A_lo, B_hi), (A_hi, B_lo)
            const int seg = kb / p.kb_seg, off = kb - seg * p.kb_seg;
            k0 = ((seg == 1 ? p.kb_seg : 0) + off) * C::BK;
            k0b = ((seg == 2 ? p.kb_seg : 0) + off) * C::BK;
          }
          if (p.a_major == MAJOR_K) {
            load(sa, &p.tma_a, k0, row0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / C::EPR; ++c) load(sa + c * (C::BK * ROW_BYTES), &p.tma_a, row0 + c * C::EPR, k0);
          }
          if (p.b_major == MAJOR_K) {
            load(sb, &p.tma_b, k0b, col0);
          } else {
#pragma unroll
            for (int c = 0; c < C::BNL / C::EPR; ++c) load(sb + c * (C::BK * ROW_BYTES), &p.tma_b, col0 + c * C::EPR, k0b);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== UMMA issuer
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(ELEM == ELEM_BF16 ? FMT_BF16 : FMT_TF32, p.a_major, p.b_major, C::TM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = worker; item < total_items; item += n_workers, ++it) {
        const int split = item % p.split_k;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < C::BK / C::UMMA_K; ++k) {
            // K-major : advance 32 bytes inside the 128-byte swizzle row, 8-row groups 1024 B apart.
            // MN-major: advance UMMA_K k-rows (128 B each); 64-element MN chunks BK*128 B apart.
            const uint64_t adesc = (p.a_major == MAJOR_K)
                                       ? make_smem_desc(sa + k * 32, 0, 1024)
                                       : make_smem_desc(sa + k * C::UMMA_K * ROW_BYTES, C::BK * ROW_BYTES, 1024);
            const uint64_t bdesc = (p.b_major == MAJOR_K)
                                       ? make_smem_desc(sb + k * 32, 0, 1024)
                                       : make_smem_desc(sb + k * C::UMMA_K * ROW_BYTES, C::BK * ROW_BYTES, 1024);
            const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
            if constexpr (PAIR) {
              if (ELEM == ELEM_BF16)
                umma_f16_ss_2sm(tmem_d, adesc, bdesc, idesc, accum);
              else
                umma_tf32_ss_2sm(tmem_d, adesc, bdesc, idesc, accum);
            } else {
              if (ELEM == ELEM_BF16)
                umma_f16_ss(tmem_d, adesc, bdesc, idesc, accum);
              else
                umma_tf32_ss(tmem_d, adesc, bdesc, idesc, accum);
            }
          }
          // smem slot reusable once these MMAs retire (PAIR: in both CTAs)
          if constexpr (PAIR)
            umma_commit_2sm(&empty_bar[stage], 3);
          else
            umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (PAIR: of both CTAs)
        if constexpr (PAIR)
          umma_commit_2sm(&tfull_bar[acc], 3);
        else
          umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ================================================================== epilogue (EW warps)
    // Each warp owns 32 accumulator rows.  Output leaves in 128-byte row segments (64 bf16 or 32 fp32 columns):
    // the thread writes its segment into a swizzled [32 x 128 B] staging tile, one lane issues a TMA tensor
    // store (or a TMA reduce-add for split-K / gradient accumulation).  Full-line, coalesced writes; rows past
    // M and columns past N are clipped by the tensor map.
    const int quad = warp & 3;
    uint8_t* stg = stg_base + (warp - 2) * 8192;
    constexpr int NSUB = EW / 4;         // warps per lane quadrant
    const int sub = (warp - 2) >> 2;     // this warp takes the column chunks c with c % NSUB == sub
    uint32_t n_st = 0;  // stores issued by this warp (staging buffer = n_st & 1)
    auto stage_out = [&](const CUtensorMap* map, const uint32_t (&w)[32], int c0, int r0, bool reduce) {
      uint8_t* buf = stg + (n_st & 1) * 4096;
      if (n_st >= 2) {
        if (lane == 0) tma_store_wait_read<1>();  // the store that used this buffer two stores ago has read it
        __syncwarp();
      }
      const uint32_t rowb = smem_u32(buf) + lane * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((u ^ (lane & 7)) << 4)), "r"(w[4 * u]),
                     "r"(w[4 * u + 1]), "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                     : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (reduce)
          tma_reduce_add_2d(map, buf, c0, r0);
        else
          tma_store_2d(map, buf, c0, r0);
        tma_store_commit();
      }
      ++n_st;
    };
    int it = 0;
    for (int item = worker; item < total_items; item += n_workers, ++it) {
      const int tile = item / p.split_k;
      const int tn = tile % p.tiles_n, tm = tile / p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int r0 = tm * C::TM + (int)rank * BM + quad * 32;
      const int m = r0 + lane;
      const uint32_t trow = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      // The accumulator is handed back as soon as this warp's LAST chunk is in registers, not after it is processed and
      // stored: at K <= 512 the issuer waits for a free accumulator on nearly every item (ncu), and the epilogue warps in turn
      // wait for the next one -- the earlier release takes the last chunk's work out of that hand-off.
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR)
            mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
          else
            mbar_arrive(&tempty_bar[acc]);
        }
      };

      if constexpr (QKV) {
        // 64-column chunks = one attention head of q, k or v.  Packed fp32x2 math throughout: with K = 512 this epilogue, not
        // the main loop, bounds the work item (two epilogue warps per sub-partition, latency-bound instruction streams).
        const float inv_d = 1.0f / 64.0f;
        const float eps = 1.1920929e-07f;  // torch.finfo(float32).eps: nn.RMSNorm(eps=None)
        const int pos = (m < p.M) ? (m % p.L) : 0;
        const bool want_raw = p.raw_out != nullptr, want_lo = p.c_split != 0;
        // rope values of this thread's token from the 32-row-transposed copy of the table
        // ([pos / 32][j][pos % 32][4], behind the [L][64] rows): the warp's 32 rows are consecutive positions, so
        // each 16-byte load is lane-consecutive (4 wavefronts per request; the row layout needed 32)
        const float4* cs = reinterpret_cast<const float4*>(p.rope + (size_t)p.L * 64 + (size_t)(pos >> 5) * 2048) + (pos & 31);
#pragma unroll 1
        for (int c = sub; c < BN / 64; c += NSUB) {
          const int n0 = tn * BN + c * 64;
          uint32_t r0v[32], r1v[32];
          __syncwarp();
          tmem_ld32(trow + c * 64, r0v);
          tmem_ld32(trow + c * 64 + 32, r1v);
          if (n0 >= p.N) {  // warp-uniform
            tmem_wait_ld();
            if (c + NSUB >= BN / 64) release_acc();
            continue;
          }
          float2 xa[16], xb[16];  // columns [0, 32) and [32, 64) of the head, two per element
          {
            float4 b0[8], b1[8];  // in flight while the TMEM load lands
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              b0[j4] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j4);
              b1[j4] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 32) + j4);
            }
            tmem_wait_ld();
            if (c + NSUB >= BN / 64) release_acc();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              xa[2 * j4] = fadd2(make_float2(__uint_as_float(r0v[4 * j4]), __uint_as_float(r0v[4 * j4 + 1])), make_float2(b0[j4].x, b0[j4].y));
              xa[2 * j4 + 1] = fadd2(make_float2(__uint_as_float(r0v[4 * j4 + 2]), __uint_as_float(r0v[4 * j4 + 3])), make_float2(b0[j4].z, b0[j4].w));
              xb[2 * j4] = fadd2(make_float2(__uint_as_float(r1v[4 * j4]), __uint_as_float(r1v[4 * j4 + 1])), make_float2(b1[j4].x, b1[j4].y));
              xb[2 * j4 + 1] = fadd2(make_float2(__uint_as_float(r1v[4 * j4 + 2]), __uint_as_float(r1v[4 * j4 + 3])), make_float2(b1[j4].z, b1[j4].w));
            }
          }
          const int which = n0 / p.dh;
          uint32_t w[32];
          if (want_raw) {
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = pack_bf16(xa[i].x, xa[i].y), w[16 + i] = pack_bf16(xb[i].x, xb[i].y);
            stage_out(&p.tma_raw, w, n0, r0, false);
          }
          if (which < 2) {
            float2 s2 = make_float2(0.f, 0.f), t2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) s2 = ffma2(xa[i], xa[i], s2), t2 = ffma2(xb[i], xb[i], t2);
            const float r = rsqrtf(((s2.x + s2.y) + (t2.x + t2.y)) * inv_d + eps);
            const float2 rr = make_float2(r, r);
            const float4* wn = reinterpret_cast<const float4*>(which == 0 ? p.qnorm_w : p.knorm_w);
            constexpr int cstep = 32;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 cc = __ldg(cs + j4 * cstep);
              const float4 sn = __ldg(cs + (8 + j4) * cstep);
              const float4 wa = __ldg(wn + j4), wb = __ldg(wn + 8 + j4);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = 2 * j4 + h;
                const float2 c2 = h ? make_float2(cc.z, cc.w) : make_float2(cc.x, cc.y);
                const float2 sv = h ? make_float2(sn.z, sn.w) : make_float2(sn.x, sn.y);
                const float2 a = fmul2(fmul2(xa[i], rr), h ? make_float2(wa.z, wa.w) : make_float2(wa.x, wa.y));
                const float2 b = fmul2(fmul2(xb[i], rr), h ? make_float2(wb.z, wb.w) : make_float2(wb.x, wb.y));
                xa[i] = ffma2(a, c2, fmul2(b, make_float2(-sv.x, -sv.y)));
                xb[i] = ffma2(a, sv, fmul2(b, c2));
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack_bf16(xa[i].x, xa[i].y), w[16 + i] = pack_bf16(xb[i].x, xb[i].y);
          stage_out(&p.tma_c, w, n0, r0, false);
          if (want_lo) {  // lo = bf16(x - float(hi))
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __nv_bfloat162 ha = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
              const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&w[16 + i]);
              w[i] = pack_bf16(xa[i].x - __low2float(ha), xa[i].y - __high2float(ha));
              w[16 + i] = pack_bf16(xb[i].x - __low2float(hb), xb[i].y - __high2float(hb));
            }
            stage_out(&p.tma_c, w, p.N + n0, r0, false);
          }
        }
      } else if (p.c_fp32) {
        // fp32 output (or fp32 reduce-add): 32 columns = 128 bytes per row segment
#pragma unroll 1
        for (int c = sub; c < BN / 32; c += NSUB) {
          const int n0 = tn * BN + c * 32;
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(trow + c * 32, r);
          tmem_wait_ld();
          if (c + NSUB >= BN / 32) release_acc();
          if (n0 >= p.N) continue;  // warp-uniform
          if (p.bias != nullptr && p.epi != EPI_ATOMIC) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j4);
              r[4 * j4] = __float_as_uint(__uint_as_float(r[4 * j4]) + b.x);
              r[4 * j4 + 1] = __float_as_uint(__uint_as_float(r[4 * j4 + 1]) + b.y);
              r[4 * j4 + 2] = __float_as_uint(__uint_as_float(r[4 * j4 + 2]) + b.z);
              r[4 * j4 + 3] = __float_as_uint(__uint_as_float(r[4 * j4 + 3]) + b.w);
            }
          }
          if (p.epi == EPI_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(silu_f(__uint_as_float(r[j])));
          }
          stage_out(&p.tma_c, r, n0, r0, p.epi == EPI_ATOMIC);
        }
      } else if (p.epi == EPI_STORE && !p.c_split && p.N % 64 == 0) {
        // bf16 output, the plain case (every bf16 GEMM of the training step): no run-time flags inside the chunk, bias loads
        // in flight while the TMEM load lands, packed adds -- a quarter of the general path's instructions, which bounded
        // the K = 512 shapes (ncu: the issuer waited for a free accumulator on 97 % of its items)
        const bool has_bias = p.bias != nullptr;
#pragma unroll 1
        for (int c = sub; c < BN / 64; c += NSUB) {
          const int n0 = tn * BN + c * 64;
          uint32_t ra[32], rb[32];
          __syncwarp();
          tmem_ld32(trow + c * 64, ra);
          tmem_ld32(trow + c * 64 + 32, rb);
          if (n0 >= p.N) {  // warp-uniform
            tmem_wait_ld();
            if (c + NSUB >= BN / 64) release_acc();
            continue;
          }
          uint32_t w[32];
          if (has_bias) {
            float4 bv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
            tmem_wait_ld();
            if (c + NSUB >= BN / 64) release_acc();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float2 a0 = fadd2(make_float2(__uint_as_float(ra[4 * j4]), __uint_as_float(ra[4 * j4 + 1])), make_float2(bv[j4].x, bv[j4].y));
              const float2 a1 = fadd2(make_float2(__uint_as_float(ra[4 * j4 + 2]), __uint_as_float(ra[4 * j4 + 3])), make_float2(bv[j4].z, bv[j4].w));
              const float2 b0 = fadd2(make_float2(__uint_as_float(rb[4 * j4]), __uint_as_float(rb[4 * j4 + 1])), make_float2(bv[8 + j4].x, bv[8 + j4].y));
              const float2 b1 = fadd2(make_float2(__uint_as_float(rb[4 * j4 + 2]), __uint_as_float(rb[4 * j4 + 3])), make_float2(bv[8 + j4].z, bv[8 + j4].w));
              w[2 * j4] = pack_bf16(a0.x, a0.y), w[2 * j4 + 1] = pack_bf16(a1.x, a1.y);
              w[16 + 2 * j4] = pack_bf16(b0.x, b0.y), w[16 + 2 * j4 + 1] = pack_bf16(b1.x, b1.y);
            }
          } else {
            tmem_wait_ld();
            if (c + NSUB >= BN / 64) release_acc();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              w[j] = pack_bf16(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
              w[16 + j] = pack_bf16(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
            }
          }
          stage_out(&p.tma_c, w, n0, r0, false);
        }
      } else {
        // bf16 output, general: 64 columns = 128 bytes per row segment (SiLU, (hi | lo) output, N % 64 == 32)
#pragma unroll 1
        for (int c = sub; c < BN / 64; c += NSUB) {
          const int n0 = tn * BN + c * 64;
          uint32_t ra[32], rb[32];
          __syncwarp();
          tmem_ld32(trow + c * 64, ra);
          tmem_ld32(trow + c * 64 + 32, rb);
          tmem_wait_ld();
          if (c + NSUB >= BN / 64) release_acc();
          if (n0 >= p.N) continue;  // warp-uniform
          uint32_t w[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t* rr = half == 0 ? ra : rb;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias != nullptr && n0 + half * 32 < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + half * 32) + j4);
              float v0 = __uint_as_float(rr[4 * j4]) + b.x, v1 = __uint_as_float(rr[4 * j4 + 1]) + b.y;
              float v2 = __uint_as_float(rr[4 * j4 + 2]) + b.z, v3 = __uint_as_float(rr[4 * j4 + 3]) + b.w;
              if (p.epi == EPI_SILU) v0 = silu_f(v0), v1 = silu_f(v1), v2 = silu_f(v2), v3 = silu_f(v3);
              w[half * 16 + 2 * j4] = pack_bf16(v0, v1);
              w[half * 16 + 2 * j4 + 1] = pack_bf16(v2, v3);
              if (p.c_split) {  // keep the residuals in place of the accumulators for the (lo) store below
                const __nv_bfloat162 ha = *reinterpret_cast<const __nv_bfloat162*>(&w[half * 16 + 2 * j4]);
                const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&w[half * 16 + 2 * j4 + 1]);
                uint32_t* rw = half == 0 ? ra : rb;
                rw[4 * j4] = __float_as_uint(v0 - __low2float(ha)), rw[4 * j4 + 1] = __float_as_uint(v1 - __high2float(ha));
                rw[4 * j4 + 2] = __float_as_uint(v2 - __low2float(hb)), rw[4 * j4 + 3] = __float_as_uint(v3 - __high2float(hb));
              }
            }
          }
          stage_out(&p.tma_c, w, n0, r0, false);
          if (p.c_split) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              w[j] = pack_bf16(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
              w[16 + j] = pack_bf16(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
            }
            stage_out(&p.tma_c, w, p.N + n0, r0, false);
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();  // all bulk stores of this warp complete before the CTA retires its smem
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (PAIR)
    cluster_sync_all();  // neither CTA retires while the other can still read its shared memory or signal its barriers
  else
    __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    if constexpr (PAIR)
      tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else
      tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
template <int BN, int ELEM, bool QKV, int EW, bool PAIR = false>
static int launch_cfg(const GemmArgs& a, cudaStream_t stream) {
  using C = Cfg<BN, ELEM, EW, PAIR>;
  GemmParams p;
  const int eb = (ELEM == ELEM_BF16) ? 2 : 4;
  // A operand
  const uint64_t Kphys = a.split3 ? 2 * (uint64_t)a.K : (uint64_t)a.K;
  if (a.a_major == MAJOR_K)
    OSD_TRY(make_tmap_2d(&p.tma_a, a.A, eb, Kphys, (uint64_t)a.M, (uint64_t)a.lda * eb, C::EPR, BM));
  else
    OSD_TRY(make_tmap_2d(&p.tma_a, a.A, eb, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda * eb, C::EPR, C::BK));
  if (a.b_major == MAJOR_K)
    OSD_TRY(make_tmap_2d(&p.tma_b, a.B, eb, Kphys, (uint64_t)a.N, (uint64_t)a.ldb * eb, C::EPR, C::BNL));
  else
    OSD_TRY(make_tmap_2d(&p.tma_b, a.B, eb, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb * eb, C::EPR, C::BK));
  {
    const int ce = a.c_fp32 ? 4 : 2;
    OSD_TRY(make_tmap_2d(&p.tma_c, a.C, ce, (uint64_t)a.N * (a.c_split ? 2 : 1), (uint64_t)a.M, (uint64_t)a.ldc * ce,
                         128 / ce, 32));
    if (a.raw_out != nullptr)
      OSD_TRY(make_tmap_2d(&p.tma_raw, a.raw_out, 2, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldc * 2, 64, 32));
    else
      p.tma_raw = p.tma_c;
  }
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.a_major = a.a_major;
  p.b_major = a.b_major;
  p.kb_seg = ceil_div(a.K, C::BK);
  p.split3 = a.split3;
  p.c_split = a.c_split;
  p.num_kb = a.split3 ? 3 * p.kb_seg : p.kb_seg;
  int split = a.split_k < 1 ? 1 : a.split_k;
  if (split > p.num_kb) split = p.num_kb;
  p.kb_per_split = ceil_div(p.num_kb, split);
  p.split_k = ceil_div(p.num_kb, p.kb_per_split);
  p.tiles_m = ceil_div(a.M, C::TM);
  p.tiles_n = ceil_div(a.N, BN);
  p.epi = a.epi;
  p.C = a.C;
  p.ldc = a.ldc;
  p.c_fp32 = a.c_fp32;
  p.bias = a.bias;
  p.qnorm_w = a.qnorm_w;
  p.knorm_w = a.knorm_w;
  p.rope = a.rope;
  p.L = a.L;
  p.dh = a.dh;
  p.raw_out = a.raw_out;

  auto kern = gemm_kernel<BN, ELEM, QKV, EW, PAIR>;
  static DeviceOnce once;  // per instantiation
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  const int total = p.tiles_m * p.tiles_n * p.split_k;
  if constexpr (PAIR) {
    const int pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (total < pairs ? total : pairs));
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    OSD_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    const int grid = total < num_sms() ? total : num_sms();
    kern<<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(p);
  }
  OSD_LAUNCHED();
  return 0;
}

// CTA pairs ([256 x 256] tiles, cta_group::2): measured at the layer's shapes (tools/gemm_ab.py, T = 131072 tokens) they win
// 6-10 % for K >= 512 (the single-CTA kernel's 96 B/clk/SM of operand loads becomes 64; ncu: tensor pipe 93 % busy at
// K = 2816) and lose at K = 128, where the epilogue bounds the item and the pair adds a cross-CTA accumulator hand-off; the
// fused QKV epilogue (K = 512, epilogue / write bound) is the same either way and stays on single CTAs.
// OSD_GEMM_PAIR=0 switches them off, =1 forces them for every eligible shape (A/B).
static int pair_mode() {
  static const int m = [] {
    const char* e = getenv("OSD_GEMM_PAIR");
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  return m;
}
// 256-wide tiles also where N is not a multiple of 256 if at most an eighth of the last tile column is padding (N = 1408: the
// BN = 128 kernel loads 128 B/clk/SM of operands at full tensor rate, 871 TFLOP/s measured; six 256-wide tiles waste 8 %)
static bool wide_n(int N) {
  const int n_wide = ceil_div(N, 256) * 256;
  return (n_wide - N) * 8 <= n_wide;
}
static bool pair_wanted(int elem, int M, int N, int kb_item, int split) {
  if (pair_mode() == 0 || elem != ELEM_BF16 || !wide_n(N)) return false;
  if (ceil_div(M, 2 * BM) * ceil_div(N, 256) * split * 10 < (num_sms() / 2) * 9) return false;  // too few pair tiles: 148 single CTAs fill better
  return pair_mode() == 1 || kb_item >= 8;
}

int gemm_split_for(int M, int N, int K) {
  const int kb = ceil_div(K, 64);
  int best = 1;
  long long best_cost = -1;
  for (int pass = 0; pass < 2; ++pass) {  // pass 0: as CTA pairs if launch_gemm would pick them for the resulting split
    const bool pair = pass == 0;
    if (pair && (pair_mode() == 0 || !wide_n(N))) continue;
    const int bn = wide_n(N) ? 256 : 128;
    const int tiles = ceil_div(M, pair ? 2 * BM : BM) * ceil_div(N, bn);
    const int workers = pair ? num_sms() / 2 : num_sms();
    const int s_max = ceil_div(8 * workers, tiles);
    for (int s = 1; s <= s_max && s <= kb; ++s) {
      const int per = ceil_div(kb, s);
      if (pair != pair_wanted(ELEM_BF16, M, N, per, s)) continue;
      const long long waves = ceil_div(tiles * s, workers);
      const long long cost = waves * (per + 6);  // + pipeline fill / epilogue of an item, in k-block times
      if (best_cost < 0 || cost < best_cost) best_cost = cost, best = s;
    }
  }
  return best;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  OSD_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem %d x %d x %d", a.M, a.N, a.K);
  OSD_CHECK(a.A && a.B && a.C, "gemm: null operand");
  OSD_CHECK(a.N % 32 == 0, "gemm: N=%d must be a multiple of 32", a.N);
  OSD_CHECK(a.epi != EPI_ATOMIC || a.c_fp32, "gemm: EPI_ATOMIC needs an fp32 output");
  OSD_CHECK(a.split_k <= 1 || a.epi == EPI_ATOMIC, "gemm: split-K needs EPI_ATOMIC");
  OSD_CHECK(!a.split3 || (a.a_major == MAJOR_K && a.b_major == MAJOR_K && a.K % (a.elem == ELEM_BF16 ? 64 : 32) == 0 &&
                          a.split_k <= 1),
            "gemm: split3 needs K-major operands with K a multiple of the stage depth (64 bf16 / 32 tf32)");
  OSD_CHECK(!a.c_split || (!a.c_fp32 && a.epi != EPI_ATOMIC && a.N % 64 == 0), "gemm: c_split needs a bf16 store epilogue");
  const int align = a.c_fp32 ? 4 : 8;
  OSD_CHECK(a.ldc % align == 0, "gemm: ldc=%lld must be a multiple of %d", (long long)a.ldc, align);
  const int kphys = (a.split3 ? 3 : 1) * ceil_div(a.K, 64);
  const int kb_item = ceil_div(kphys, a.split_k < 1 ? 1 : a.split_k);
  const bool pair = pair_wanted(a.elem, a.M, a.N, kb_item, a.split_k < 1 ? 1 : a.split_k) && (a.epi != EPI_QKV || pair_mode() == 1);
  if (a.epi == EPI_QKV) {
    OSD_CHECK(a.elem == ELEM_BF16 || a.elem == ELEM_TF32, "gemm: bad elem");
    OSD_CHECK(a.N % 64 == 0 && a.dh % 64 == 0 && a.bias && a.qnorm_w && a.knorm_w && a.rope && a.L > 0 && !a.c_fp32,
              "gemm: EPI_QKV arguments incomplete");
    if (a.elem == ELEM_BF16 && pair) return launch_cfg<256, ELEM_BF16, true, 8, true>(a, stream);
    if (a.elem == ELEM_BF16) return launch_cfg<256, ELEM_BF16, true, 8>(a, stream);
    return launch_cfg<256, ELEM_TF32, true, 4>(a, stream);
  }
  // narrow tiles when the wide ones cannot fill the machine or would be more than an eighth padding (wide_n)
  const int wide_items = ceil_div(a.M, BM) * ceil_div(a.N, 256) * (a.split_k < 1 ? 1 : a.split_k);
  const bool narrow = !wide_n(a.N) || wide_items < num_sms();
  // k-blocks per work item: short main loops cannot hide a 4-warp epilogue (OSD_GEMM_EW=4 forces the old layout)
  static const bool ew4_only = [] {
    const char* e = getenv("OSD_GEMM_EW");
    return e != nullptr && e[0] == '4';
  }();
  const bool wide_epi = kb_item <= 8 && !ew4_only;  // K <= 512: measured faster with 8 (tools/gemm_ab.py); K >= 1024 slower
  if (a.elem == ELEM_BF16) {
    if (pair)  // pair_wanted: enough [256 x 256] tiles for (nearly) every CTA pair
      return wide_epi ? launch_cfg<256, ELEM_BF16, false, 8, true>(a, stream) : launch_cfg<256, ELEM_BF16, false, 4, true>(a, stream);
    if (narrow) return wide_epi ? launch_cfg<128, ELEM_BF16, false, 8>(a, stream) : launch_cfg<128, ELEM_BF16, false, 4>(a, stream);
    return wide_epi ? launch_cfg<256, ELEM_BF16, false, 8>(a, stream) : launch_cfg<256, ELEM_BF16, false, 4>(a, stream);
  } else {
    if (narrow) return launch_cfg<128, ELEM_TF32, false, 4>(a, stream);
    return launch_cfg<256, ELEM_TF32, false, 4>(a, stream);
  }
}

}  // namespace osd
