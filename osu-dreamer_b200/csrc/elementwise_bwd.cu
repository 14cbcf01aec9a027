// Backward of the HBM-bound kernels in elementwise.cu.  Same token-major layout; per-channel parameter /
// modulation gradients are reduced inside the block (registers -> shared atomics) before one global
// atomic per channel per block.  Blocks never straddle two samples (grid.y = batch).
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

static constexpr int D = 512;
static constexpr float RMS_EPS = 1e-6f;
static constexpr int TOKB = 64;  // tokens per block for the row-wise kernels (8 warps x 8 rows)

__device__ __forceinline__ void ld_row_f32(const float* __restrict__ p, float (&r)[16], int lane) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 q = *reinterpret_cast<const float4*>(p + v * 128 + lane * 4);
    r[4 * v] = q.x, r[4 * v + 1] = q.y, r[4 * v + 2] = q.z, r[4 * v + 3] = q.w;
  }
}
__device__ __forceinline__ void ld_row_bf16(const __nv_bfloat16* __restrict__ p, float (&r)[16], int lane) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const uint2 q = *reinterpret_cast<const uint2*>(p + v * 128 + lane * 4);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
    r[4 * v] = __low2float(a), r[4 * v + 1] = __high2float(a), r[4 * v + 2] = __low2float(b), r[4 * v + 3] = __high2float(b);
  }
}
__device__ __forceinline__ void st_row_f32(float* __restrict__ p, const float (&r)[16], int lane) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
    *reinterpret_cast<float4*>(p + v * 128 + lane * 4) = make_float4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
}
__device__ __forceinline__ void st_row_bf16(__nv_bfloat16* __restrict__ p, const float (&r)[16], int lane) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
    *reinterpret_cast<uint2*>(p + v * 128 + lane * 4) =
        make_uint2(pack_bf16(r[4 * v], r[4 * v + 1]), pack_bf16(r[4 * v + 2], r[4 * v + 3]));
}
__device__ __forceinline__ float inv_rms16(const float (&r)[16]) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) ss = fmaf(r[i], r[i], ss);
  return rsqrtf(warp_sum(ss) * (1.0f / D) + RMS_EPS);
}
// lane-held per-channel partials (16 per lane) -> smem[512] (shared atomics) ; call flush_smem512 after a sync
__device__ __forceinline__ void acc_to_smem(float* sm, const float (&a)[16], int lane) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(sm + v * 128 + lane * 4 + i, a[4 * v + i]);
}

// ------------------------------------------------------------------------------------------------
// final rms_norm + proj_out backward: dx = d(rms_norm)(Wout^T dv), dWout, dbout.
__global__ void __launch_bounds__(256) final_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dv,
                                                        const float* __restrict__ Wo, float* __restrict__ dx,
                                                        float* __restrict__ dWo, float* __restrict__ dbo, int L) {
  __shared__ float sW[6 * D];
  __shared__ float sWo[6 * D];
  __shared__ float sB[8];
  const int b = blockIdx.y, l0 = blockIdx.x * TOKB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * D; i += 256) sW[i] = 0.f, sWo[i] = Wo[i];
  if (threadIdx.x < 8) sB[threadIdx.x] = 0.f;
  __syncthreads();
  float aw[6][16];
#pragma unroll
  for (int e = 0; e < 6; ++e)
#pragma unroll
    for (int i = 0; i < 16; ++i) aw[e][i] = 0.f;
  float ab[6] = {0, 0, 0, 0, 0, 0};
  for (int rr = warp; rr < TOKB; rr += 8) {
    const int l = l0 + rr;
    if (l >= L) break;
    const size_t t = (size_t)b * L + l;
    float xr[16];
    ld_row_f32(x + t * D, xr, lane);
    const float inv = inv_rms16(xr);
    float g[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) g[e] = __ldg(dv + ((size_t)b * 6 + e) * L + l);
    float dn[16];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float n = xr[i] * inv;
      const int col = (i >> 2) * 128 + lane * 4 + (i & 3);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        d = fmaf(sWo[e * D + col], g[e], d);
        aw[e][i] = fmaf(g[e], n, aw[e][i]);
      }
      dn[i] = d;
      dot = fmaf(d, n, dot);
      xr[i] = n;
    }
    dot = warp_sum(dot) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < 16; ++i) dn[i] = inv * (dn[i] - xr[i] * dot);
    st_row_f32(dx + t * D, dn, lane);
#pragma unroll
    for (int e = 0; e < 6; ++e) ab[e] += g[e];
  }
#pragma unroll
  for (int e = 0; e < 6; ++e) acc_to_smem(sW + e * D, aw[e], lane);
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < 6; ++e) atomicAdd(&sB[e], ab[e]);
  __syncthreads();
  for (int i = threadIdx.x; i < 6 * D; i += 256) atomicAdd(dWo + i, sW[i]);
  if (threadIdx.x < 6) atomicAdd(dbo + threadIdx.x, sB[threadIdx.x]);
}
int launch_final_bwd(const float* x, const float* dv, const float* Wo, float* dx, float* dWo, float* dbo, int B, int L,
                     cudaStream_t s) {
  dim3 grid(ceil_div(L, TOKB), B);
  final_bwd_kernel<<<grid, 256, 0, s>>>(x, dv, Wo, dx, dWo, dbo, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// gated residual backward: x_out = x + rms_norm(h) * gate  ->  dh (bf16 operand), dgate[b], dbias = colsum(dh).
// dx passes through unchanged (same buffer).
__global__ void __launch_bounds__(256) postnorm_gate_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ h,
                                                                const float* __restrict__ mod,
                                                                __nv_bfloat16* __restrict__ dh, float* __restrict__ dmod,
                                                                float* __restrict__ dbias, int L) {
  __shared__ float sG[D], sBias[D];
  const int b = blockIdx.y, l0 = blockIdx.x * TOKB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < D; i += 256) sG[i] = 0.f, sBias[i] = 0.f;
  __syncthreads();
  float g[16];
  ld_row_f32(mod + (size_t)b * 1536 + 1024, g, lane);
  float ag[16], abias[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ag[i] = 0.f, abias[i] = 0.f;
  for (int rr = warp; rr < TOKB; rr += 8) {
    const int l = l0 + rr;
    if (l >= L) break;
    const size_t t = (size_t)b * L + l;
    float hr[16], dr[16];
    ld_row_f32(h + t * D, hr, lane);
    ld_row_f32(dx + t * D, dr, lane);
    const float inv = inv_rms16(hr);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float n = hr[i] * inv;
      ag[i] = fmaf(dr[i], n, ag[i]);
      dr[i] *= g[i];  // dn
      dot = fmaf(dr[i], n, dot);
      hr[i] = n;
    }
    dot = warp_sum(dot) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      dr[i] = inv * (dr[i] - hr[i] * dot);
      abias[i] += dr[i];
    }
    st_row_bf16(dh + t * D, dr, lane);
  }
  acc_to_smem(sG, ag, lane);
  acc_to_smem(sBias, abias, lane);
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += 256) {
    atomicAdd(dmod + (size_t)b * 1536 + 1024 + i, sG[i]);
    atomicAdd(dbias + i, sBias[i]);
  }
}
int launch_postnorm_gate_bwd(const float* dx, const float* h, const float* mod, void* dh, float* dmod, float* dbias,
                             int B, int L, cudaStream_t s) {
  dim3 grid(ceil_div(L, TOKB), B);
  postnorm_gate_bwd_kernel<<<grid, 256, 0, s>>>(dx, h, mod, static_cast<__nv_bfloat16*>(dh), dmod, dbias, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// attn front end backward: z = rms_norm(x)*(1+scale)+shift + cl
//   dx += d(rms_norm)(dz*(1+scale)); dscale[b], dshift[b]; dbcl = colsum(dz)   (dcl = dz feeds the proj_cl GEMMs)
__global__ void __launch_bounds__(256) prenorm_mod_bwd_kernel(const __nv_bfloat16* __restrict__ dz,
                                                              const float* __restrict__ x, const float* __restrict__ mod,
                                                              float* __restrict__ dx, float* __restrict__ dmod,
                                                              float* __restrict__ dbcl, int L) {
  __shared__ float sSc[D], sSh[D];
  const int b = blockIdx.y, l0 = blockIdx.x * TOKB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < D; i += 256) sSc[i] = 0.f, sSh[i] = 0.f;
  __syncthreads();
  float sc[16];
  ld_row_f32(mod + (size_t)b * 1536, sc, lane);
  float asc[16], ash[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) asc[i] = 0.f, ash[i] = 0.f;
  for (int rr = warp; rr < TOKB; rr += 8) {
    const int l = l0 + rr;
    if (l >= L) break;
    const size_t t = (size_t)b * L + l;
    float xr[16], g[16], dr[16];
    ld_row_f32(x + t * D, xr, lane);
    ld_row_bf16(dz + t * D, g, lane);
    ld_row_f32(dx + t * D, dr, lane);
    const float inv = inv_rms16(xr);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float n = xr[i] * inv;
      asc[i] = fmaf(g[i], n, asc[i]);
      ash[i] += g[i];
      g[i] *= (1.f + sc[i]);  // dn
      dot = fmaf(g[i], n, dot);
      xr[i] = n;
    }
    dot = warp_sum(dot) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < 16; ++i) dr[i] += inv * (g[i] - xr[i] * dot);
    st_row_f32(dx + t * D, dr, lane);
  }
  acc_to_smem(sSc, asc, lane);
  acc_to_smem(sSh, ash, lane);
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += 256) {
    atomicAdd(dmod + (size_t)b * 1536 + i, sSc[i]);
    atomicAdd(dmod + (size_t)b * 1536 + 512 + i, sSh[i]);
    if (dbcl != nullptr) atomicAdd(dbcl + i, sSh[i]);
  }
}
int launch_prenorm_mod_bwd(const void* dz, const float* x, const float* mod, float* dx, float* dmod, float* dbcl, int B,
                           int L, cudaStream_t s) {
  dim3 grid(ceil_div(L, TOKB), B);
  prenorm_mod_bwd_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dz), x, mod, dx, dmod, dbcl, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ffn front end backward: z2 = dwconv5(hmod) + b,  hmod = rms_norm(x1)*(1+scale)+shift
//   dhmod[l] = sum_k w[k] dz2[l-k+2];  dw[k] = sum_l dz2[l] hmod[l+k-2];  db = sum dz2
//   dx += d(rms_norm)(dhmod*(1+scale));  dscale[b], dshift[b]
static constexpr int DWB_TOK = 32;   // tokens per sub-tile
static constexpr int DWB_SUB = 4;    // sub-tiles per block
__global__ void __launch_bounds__(256) dwconv_prenorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ dz2, const __nv_bfloat16* __restrict__ hmod, const float* __restrict__ x1,
    const float* __restrict__ mod, const float* __restrict__ wconv, float* __restrict__ dx, float* __restrict__ dmod,
    float* __restrict__ dw, float* __restrict__ db, int L) {
  extern __shared__ uint8_t dsm[];
  __nv_bfloat16* sG = reinterpret_cast<__nv_bfloat16*>(dsm);                      // dz2  [36][512]
  __nv_bfloat16* sH = sG + (DWB_TOK + 4) * D;                                     // hmod [36][512]
  float* sSc = reinterpret_cast<float*>(sH + (DWB_TOK + 4) * D);                  // [512]
  float* sSh = sSc + D;
  float* sWk = sSh + D;                                                           // taps transposed [5][512]
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < D; i += 256) {
    sSc[i] = 0.f, sSh[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) sWk[k * D + i] = wconv[i * 5 + k];
  }
  float sc[16];
  ld_row_f32(mod + (size_t)b * 1536, sc, lane);
  float asc[16], ash[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) asc[i] = 0.f, ash[i] = 0.f;
  // channel-mapped accumulators (thread owns channels threadIdx.x and threadIdx.x + 256)
  float adw[2][5], adb[2] = {0.f, 0.f};
#pragma unroll
  for (int cc = 0; cc < 2; ++cc)
#pragma unroll
    for (int k = 0; k < 5; ++k) adw[cc][k] = 0.f;

  for (int sub = 0; sub < DWB_SUB; ++sub) {
    const int l0 = (blockIdx.x * DWB_SUB + sub) * DWB_TOK;
    if (l0 >= L) break;
    __syncthreads();
    // phase 1: stage dz2 and hmod rows l0-2 .. l0+33 (zeros outside the sample)
    for (int rr = warp; rr < DWB_TOK + 4; rr += 8) {
      const int l = l0 + rr - 2;
      uint2 g[4], hh[4];
      if (l >= 0 && l < L) {
        const size_t t = (size_t)b * L + l;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          g[v] = *reinterpret_cast<const uint2*>(dz2 + t * D + v * 128 + lane * 4);
          hh[v] = *reinterpret_cast<const uint2*>(hmod + t * D + v * 128 + lane * 4);
        }
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) g[v] = make_uint2(0, 0), hh[v] = make_uint2(0, 0);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        *reinterpret_cast<uint2*>(sG + rr * D + v * 128 + lane * 4) = g[v];
        *reinterpret_cast<uint2*>(sH + rr * D + v * 128 + lane * 4) = hh[v];
      }
    }
    __syncthreads();
    // phase 2 (row mapped): dhmod -> dx, dscale, dshift
    for (int r = warp; r < DWB_TOK; r += 8) {
      const int l = l0 + r;
      if (l >= L) break;
      const size_t t = (size_t)b * L + l;
      float dh[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) dh[i] = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float g[16], wt[16];
        ld_row_bf16(sG + (r + 4 - k) * D, g, lane);
        ld_row_f32(sWk + k * D, wt, lane);
#pragma unroll
        for (int i = 0; i < 16; ++i) dh[i] = fmaf(wt[i], g[i], dh[i]);
      }
      float xr[16], dr[16];
      ld_row_f32(x1 + t * D, xr, lane);
      ld_row_f32(dx + t * D, dr, lane);
      const float inv = inv_rms16(xr);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float n = xr[i] * inv;
        asc[i] = fmaf(dh[i], n, asc[i]);
        ash[i] += dh[i];
        dh[i] *= (1.f + sc[i]);
        dot = fmaf(dh[i], n, dot);
        xr[i] = n;
      }
      dot = warp_sum(dot) * (1.0f / D);
#pragma unroll
      for (int i = 0; i < 16; ++i) dr[i] += inv * (dh[i] - xr[i] * dot);
      st_row_f32(dx + t * D, dr, lane);
    }
    // phase 3 (channel mapped): dw, db over this sub-tile's own output rows
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = threadIdx.x + cc * 256;
      for (int r = 0; r < DWB_TOK; ++r) {
        if (l0 + r >= L) break;
        const float g = __bfloat162float(sG[(r + 2) * D + c]);
        adb[cc] += g;
#pragma unroll
        for (int k = 0; k < 5; ++k) adw[cc][k] = fmaf(g, __bfloat162float(sH[(r + k) * D + c]), adw[cc][k]);
      }
    }
  }
  acc_to_smem(sSc, asc, lane);
  acc_to_smem(sSh, ash, lane);
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += 256) {
    atomicAdd(dmod + (size_t)b * 1536 + i, sSc[i]);
    atomicAdd(dmod + (size_t)b * 1536 + 512 + i, sSh[i]);
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = threadIdx.x + cc * 256;
    atomicAdd(db + c, adb[cc]);
#pragma unroll
    for (int k = 0; k < 5; ++k) atomicAdd(dw + c * 5 + k, adw[cc][k]);
  }
}
int launch_dwconv_prenorm_bwd(const void* dz2, const void* hmod, const float* x1, const float* mod, const float* wconv,
                              float* dx, float* dmod, float* dw, float* db, int B, int L, cudaStream_t s) {
  const int smem = 2 * (DWB_TOK + 4) * D * 2 + 7 * D * 4;
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(dwconv_prenorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid(ceil_div(L, DWB_TOK * DWB_SUB), B);
  dwconv_prenorm_bwd_kernel<<<grid, 256, smem, s>>>(static_cast<const __nv_bfloat16*>(dz2),
                                                    static_cast<const __nv_bfloat16*>(hmod), x1, mod, wconv, dx, dmod,
                                                    dw, db, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SwiGLU + RMSNorm(1365) backward: dvg from dhn; dbvg (padded layout) = column sums of dvg (fused: the bf16
// values the GEMMs read are summed here, no second pass over dvg).
// Column-parallel: a block of 11 warps owns SWB_ROWS consecutive tokens and all 2 x 1408 columns; each lane keeps
// the same 4 v and 4 g columns for every row, so the column sums stay in registers.  Rows go through in batches of
// SWB_R: all loads of a batch are issued up front (12 x 8 bytes per lane in flight), the row statistic
// mean(dhn * hn) is a warp sum + an 11-entry exchange through shared memory (double-buffered: one barrier per
// batch), and the outputs are produced from the registers of the same single read.
static constexpr int HID = 1365, HIDP = 1408;
static constexpr int SWB_ROWS = 128, SWB_R = 4, SWB_WARPS = HIDP / 128, SWB_ST = 3;
static constexpr int SWB_VG_BYTES = SWB_R * 2 * HIDP * 2, SWB_D_BYTES = SWB_R * HIDP * 2;
static constexpr int SWB_STAGE = SWB_VG_BYTES + SWB_D_BYTES;  // 4 rows of vg (contiguous) + 4 rows of dhn (contiguous)
static constexpr int SWB_SMEM = SWB_ST * SWB_STAGE + 64;
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// Rows reach the block as 1-D bulk copies (cp.async.bulk): SWB_R consecutive rows of vg and of dhn are contiguous in
// memory, so a stage is two copies; thread 0 keeps SWB_ST stages in flight (a stage is refilled right after the block
// barrier that follows its register reads), so the bytes in flight do not depend on registers or on instruction
// scheduling.
__global__ void __launch_bounds__(SWB_WARPS * 32, 2) swiglu_norm_bwd_kernel(const __nv_bfloat16* __restrict__ vg,
                                                                         const __nv_bfloat16* __restrict__ dhn,
                                                                         const float* __restrict__ rinv,
                                                                         __nv_bfloat16* __restrict__ dvg,
                                                                         float* __restrict__ dbvg, int T) {
  extern __shared__ __align__(128) uint8_t swb[];
  __shared__ float sdot[2][SWB_R][SWB_WARPS + 1];
  uint64_t* full = reinterpret_cast<uint64_t*>(swb + SWB_ST * SWB_STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = warp * 128 + lane * 4;
  const int t0 = blockIdx.x * SWB_ROWS;
  const int t1 = min(T, t0 + SWB_ROWS);
  const int nb = (t1 - t0 + SWB_R - 1) / SWB_R;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SWB_ST; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int k) {  // batch k -> stage k % SWB_ST (thread 0 only); the last batch of a block may be short
    if (k >= nb) return;
    const int tb = t0 + k * SWB_R;
    const int rows = min(SWB_R, t1 - tb);
    uint8_t* st = swb + (k % SWB_ST) * SWB_STAGE;
    mbar_expect_tx(&full[k % SWB_ST], (uint32_t)rows * (2 * HIDP * 2 + HIDP * 2));
    bulk_load_1d(st, vg + (size_t)tb * 2 * HIDP, (uint32_t)rows * 2 * HIDP * 2, &full[k % SWB_ST]);
    bulk_load_1d(st + SWB_VG_BYTES, dhn + (size_t)tb * HIDP, (uint32_t)rows * HIDP * 2, &full[k % SWB_ST]);
  };
  if (threadIdx.x == 0) {
    for (int k = 0; k < SWB_ST; ++k) issue(k);
  }
  float csv[4] = {0.f, 0.f, 0.f, 0.f}, csg[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < nb; ++k) {
    const int tb = t0 + k * SWB_R;
    const uint8_t* st = swb + (k % SWB_ST) * SWB_STAGE;
    float rr[SWB_R];
#pragma unroll
    for (int r = 0; r < SWB_R; ++r) rr[r] = rinv[min(tb + r, t1 - 1)];
    mbar_wait(&full[k % SWB_ST], (k / SWB_ST) & 1);
    uint2 va[SWB_R], ga[SWB_R], da[SWB_R];
#pragma unroll
    for (int r = 0; r < SWB_R; ++r) {  // rows past t1 hold stale bytes of an earlier batch: finite, results discarded
      const uint8_t* row = st + r * (2 * HIDP * 2);
      va[r] = *reinterpret_cast<const uint2*>(row + c0 * 2);
      ga[r] = *reinterpret_cast<const uint2*>(row + (HIDP + c0) * 2);
      da[r] = *reinterpret_cast<const uint2*>(st + SWB_VG_BYTES + r * (HIDP * 2) + c0 * 2);
    }
    float part[SWB_R];
#pragma unroll
    for (int r = 0; r < SWB_R; ++r) {
      float v[4], g[4], d[4];
      unpack_bf16x4(va[r], v);
      unpack_bf16x4(ga[r], g);
      unpack_bf16x4(da[r], d);
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) acc = fmaf(d[e], v[e] * g[e] * sigmoid_fast(g[e]), acc);
      part[r] = acc;
    }
#pragma unroll
    for (int r = 0; r < SWB_R; ++r) part[r] = warp_sum(part[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < SWB_R; ++r) sdot[k & 1][r][warp] = part[r];
    }
    __syncthreads();  // (a) the row statistics are complete; (b) every thread has copied its part of stage k out
    if (threadIdx.x == 0) {
      fence_proxy_async_smem();
      issue(k + SWB_ST);  // refill the stage just drained: SWB_ST batches stay in flight
    }
#pragma unroll
    for (int r = 0; r < SWB_R; ++r) {
      const int t = tb + r;
      if (t >= t1) break;  // block-uniform
      float dot = 0.f;
#pragma unroll
      for (int w = 0; w < SWB_WARPS; ++w) dot += sdot[k & 1][r][w];
      const float rinv_t = rr[r];
      dot = dot * rinv_t * (1.0f / HID);  // mean(dhn * hn), hn = hs * r
      float v[4], g[4], d[4];
      unpack_bf16x4(va[r], v);
      unpack_bf16x4(ga[r], g);
      unpack_bf16x4(da[r], d);
      float ov[4], og[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float sg = sigmoid_fast(g[e]);
        const float sil = g[e] * sg;
        const float dhs = rinv_t * (d[e] - v[e] * sil * rinv_t * dot);
        ov[e] = dhs * sil;
        og[e] = dhs * v[e] * (sg * (1.0f + g[e] * (1.0f - sg)));
      }
      const uint2 pv = make_uint2(pack_bf16(ov[0], ov[1]), pack_bf16(ov[2], ov[3]));
      const uint2 pg = make_uint2(pack_bf16(og[0], og[1]), pack_bf16(og[2], og[3]));
      *reinterpret_cast<uint2*>(dvg + (size_t)t * 2 * HIDP + c0) = pv;
      *reinterpret_cast<uint2*>(dvg + (size_t)t * 2 * HIDP + HIDP + c0) = pg;
      float rv[4], rg[4];
      unpack_bf16x4(pv, rv);
      unpack_bf16x4(pg, rg);
#pragma unroll
      for (int e = 0; e < 4; ++e) csv[e] += rv[e], csg[e] += rg[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    atomicAdd(dbvg + c0 + e, csv[e]);
    atomicAdd(dbvg + HIDP + c0 + e, csg[e]);
  }
}
int launch_swiglu_norm_bwd(const void* vg, const void* dhn, const float* rinv, void* dvg, float* dbvg, int T,
                           cudaStream_t s) {
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(swiglu_norm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SWB_SMEM));
  }
  swiglu_norm_bwd_kernel<<<ceil_div(T, SWB_ROWS), SWB_WARPS * 32, SWB_SMEM, s>>>(static_cast<const __nv_bfloat16*>(vg),
                                                                        static_cast<const __nv_bfloat16*>(dhn), rinv,
                                                                        static_cast<__nv_bfloat16*>(dvg), dbvg, T);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm(64)*w + RoPE backward, in place on dqkv [T,3072] (dq | dk | dv) -> gradients of the raw
// projections; dqn_w / dkn_w [64]; dbqkv [3072] = column sums of the result (fused: no second pass over dqkv).
// Column-parallel: a block owns QKB consecutive tokens of one sample and ALL 3072 columns; warp w owns the four
// 64-wide heads 4w..4w+3 of q (w < 4) or k (w >= 4) plus 128 of the v columns, for every row of the block, so the
// column sums (and the norm-weight gradients) stay in registers until the end.  8 lanes per head: lane j holds
// elements 4j..4j+3 and 32+4j..32+4j+3 -- the rope partner pairs -- and the two 64-wide row statistics are
// 3-step shuffles.  dq can be taken straight from the fp32 accumulator of the single-pass attention backward
// (dq_acc, scaled by dq_scale) when its fallback flag is clear, which replaces that path's convert pass.
static constexpr int QKB = 128;
#ifndef OSD_QKB_BLOCKS
#define OSD_QKB_BLOCKS 2
#endif
static constexpr int QKB_BLOCKS = OSD_QKB_BLOCKS;  // resident blocks per SM the register allocation targets (A/B: 3 spills 96 B)
// gin aliases dqkv: every element is read (through gin) by the thread that later overwrites it and by no other
// thread, so declaring the read side const/restrict is safe and lets the compiler hoist the loads of the next
// rows above the stores of the current one.  (A bulk-copy staged variant of this kernel -- 3 x 48 KB stages, one block
// of 8 warps per SM -- was slower, 0.745 vs 0.516 ms: at 142 registers / 8 warps the per-row math bound it.)
__global__ void __launch_bounds__(256, QKB_BLOCKS) qknorm_rope_bwd_kernel(__nv_bfloat16* __restrict__ dqkv,
                                                              const __nv_bfloat16* __restrict__ gin,
                                                              const __nv_bfloat16* __restrict__ raw,
                                                              const float* __restrict__ rope,
                                                              const float* __restrict__ qw, const float* __restrict__ kw,
                                                              float* __restrict__ dqw, float* __restrict__ dkw,
                                                              float* __restrict__ dbias,
                                                              const float* __restrict__ dq_acc,
                                                              const int* __restrict__ dq_flag, float dq_scale, int L) {
  __shared__ float sW[2][64];
  if (threadIdx.x < 128) sW[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, l0 = blockIdx.x * QKB;
  const int l1 = min(L, l0 + QKB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, j = lane & 7;
  const int isk = warp >> 2;                        // warp-uniform
  const int col = (warp * 4 + sub) * 64 + 4 * j;    // first of this lane's 4 + 4 q/k columns (second group at +32)
  const int vcol = 2048 + warp * 128 + lane * 4;    // this lane's 4 v columns
  const bool from_acc = (isk == 0) && dq_acc != nullptr && (dq_flag == nullptr || *dq_flag == 0);
  const float eps = 1.1920929e-07f;
  const float* wsrc = isk ? kw : qw;
  const float4 w0 = *reinterpret_cast<const float4*>(wsrc + 4 * j);
  const float4 w1 = *reinterpret_cast<const float4*>(wsrc + 32 + 4 * j);
  const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float adw[8], cs[8], csv[4];
#pragma unroll
  for (int e = 0; e < 8; ++e) adw[e] = 0.f, cs[e] = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) csv[e] = 0.f;

  constexpr int R = 4;  // rows per batch: all loads of a batch are issued before any of its math
  for (int lb = l0; lb < l1; lb += R) {
    uint2 xa[R], xb[R], gv[R];
    uint4 g0[R], g1[R];  // dq / dk: 4 + 4 fp32 (from the accumulator) or 4 + 4 bf16 in .x/.y
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int l = min(lb + r, l1 - 1);  // tail rows re-read the last row (results discarded)
      const size_t t = (size_t)b * L + l;
      const __nv_bfloat16* grow = gin + t * 3072;
      const __nv_bfloat16* xrow = raw + t * 3072;
      xa[r] = *reinterpret_cast<const uint2*>(xrow + col);
      xb[r] = *reinterpret_cast<const uint2*>(xrow + col + 32);
      gv[r] = *reinterpret_cast<const uint2*>(grow + vcol);
      if (from_acc) {
        const float* arow = dq_acc + t * 1024 + col;
        g0[r] = *reinterpret_cast<const uint4*>(arow);
        g1[r] = *reinterpret_cast<const uint4*>(arow + 32);
      } else {
        const uint2 ga = *reinterpret_cast<const uint2*>(grow + col);
        const uint2 gb = *reinterpret_cast<const uint2*>(grow + col + 32);
        g0[r] = make_uint4(ga.x, ga.y, 0u, 0u);
        g1[r] = make_uint4(gb.x, gb.y, 0u, 0u);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int l = lb + r;
      if (l >= l1) break;  // block-uniform
      const size_t t = (size_t)b * L + l;
      __nv_bfloat16* drow = dqkv + t * 3072;
      const float4 cc = *reinterpret_cast<const float4*>(rope + (size_t)l * 64 + 4 * j);
      const float4 sn = *reinterpret_cast<const float4*>(rope + (size_t)l * 64 + 32 + 4 * j);
      float gy[8];
      if (from_acc) {
        gy[0] = __uint_as_float(g0[r].x) * dq_scale, gy[1] = __uint_as_float(g0[r].y) * dq_scale;
        gy[2] = __uint_as_float(g0[r].z) * dq_scale, gy[3] = __uint_as_float(g0[r].w) * dq_scale;
        gy[4] = __uint_as_float(g1[r].x) * dq_scale, gy[5] = __uint_as_float(g1[r].y) * dq_scale;
        gy[6] = __uint_as_float(g1[r].z) * dq_scale, gy[7] = __uint_as_float(g1[r].w) * dq_scale;
      } else {
        unpack_bf16x4(make_uint2(g0[r].x, g0[r].y), gy);
        unpack_bf16x4(make_uint2(g1[r].x, g1[r].y), gy + 4);
      }
      float xx[8], vv[4];
      unpack_bf16x4(xa[r], xx);
      unpack_bf16x4(xb[r], xx + 4);
      unpack_bf16x4(gv[r], vv);
#pragma unroll
      for (int e = 0; e < 4; ++e) csv[e] += vv[e];
      const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
      const float sv[4] = {sn.x, sn.y, sn.z, sn.w};
      // inverse rotation: da = g1*c + g2*s ; db = -g1*s + g2*c
      float da[8];
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        da[e] = gy[e] * cv[e] + gy[4 + e] * sv[e];
        da[4 + e] = gy[4 + e] * cv[e] - gy[e] * sv[e];
        ss = fmaf(xx[e], xx[e], ss);
        ss = fmaf(xx[4 + e], xx[4 + e], ss);
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      const float rs = rsqrtf(ss * (1.0f / 64.0f) + eps);
      float n[8], dn[8];
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        n[e] = xx[e] * rs;
        dn[e] = da[e] * wv[e];
        adw[e] = fmaf(da[e], n[e], adw[e]);
        dot = fmaf(dn[e], n[e], dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot *= (1.0f / 64.0f);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = rs * (dn[2 * e] - n[2 * e] * dot), d1 = rs * (dn[2 * e + 1] - n[2 * e + 1] * dot);
        o[e] = pack_bf16(d0, d1);
        // the bias gradient is the column sum of what the GEMMs will read, i.e. of the bf16-rounded values
        cs[2 * e] += __uint_as_float(o[e] << 16);
        cs[2 * e + 1] += __uint_as_float(o[e] & 0xffff0000u);
      }
      *reinterpret_cast<uint2*>(drow + col) = make_uint2(o[0], o[1]);
      *reinterpret_cast<uint2*>(drow + col + 32) = make_uint2(o[2], o[3]);
    }
  }
  // ---- column sums: one global atomic per column per block
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    atomicAdd(dbias + col + e, cs[e]);
    atomicAdd(dbias + col + 32 + e, cs[4 + e]);
    atomicAdd(dbias + vcol + e, csv[e]);
  }
  // ---- norm-weight gradients: sum over the warp's four heads (lanes with equal j), then over warps
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    adw[e] += __shfl_xor_sync(0xffffffffu, adw[e], 8);
    adw[e] += __shfl_xor_sync(0xffffffffu, adw[e], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(&sW[isk][4 * j + e], adw[e]);
      atomicAdd(&sW[isk][32 + 4 * j + e], adw[4 + e]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(dqw + threadIdx.x, sW[0][threadIdx.x]);
  else if (threadIdx.x < 128) atomicAdd(dkw + threadIdx.x - 64, sW[1][threadIdx.x - 64]);
}
int launch_qknorm_rope_bwd(void* dqkv, const void* raw, const float* rope, const float* qw, const float* kw,
                           float* dqw, float* dkw, float* dbias, const float* dq_acc, const int* dq_flag,
                           float dq_scale, int B, int L, cudaStream_t s) {
  dim3 grid(ceil_div(L, QKB), B);
  qknorm_rope_bwd_kernel<<<grid, 256, 0, s>>>(static_cast<__nv_bfloat16*>(dqkv), static_cast<const __nv_bfloat16*>(dqkv),
                                              static_cast<const __nv_bfloat16*>(raw),
                                              rope, qw, kw, dqw, dkw, dbias, dq_acc, dq_flag, dq_scale, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// proj_in backward: dW[c][e] = sum_t dx[t][c] xt[b][e][l], db[c] = sum_t dx[t][c]  (xt is data: no dxt).
__global__ void __launch_bounds__(256) proj_in_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ xt,
                                                          float* __restrict__ dW, float* __restrict__ db, int L) {
  __shared__ float sW[7 * D];
  for (int i = threadIdx.x; i < 7 * D; i += 256) sW[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, l0 = blockIdx.x * TOKB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[7][16];
#pragma unroll
  for (int e = 0; e < 7; ++e)
#pragma unroll
    for (int i = 0; i < 16; ++i) aw[e][i] = 0.f;
  for (int rr = warp; rr < TOKB; rr += 8) {
    const int l = l0 + rr;
    if (l >= L) break;
    const size_t t = (size_t)b * L + l;
    float g[16];
    ld_row_f32(dx + t * D, g, lane);
    float in[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) in[e] = __ldg(xt + ((size_t)b * 6 + e) * L + l);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
      for (int e = 0; e < 6; ++e) aw[e][i] = fmaf(g[i], in[e], aw[e][i]);
      aw[6][i] += g[i];
    }
  }
#pragma unroll
  for (int e = 0; e < 7; ++e) acc_to_smem(sW + e * D, aw[e], lane);
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += 256) {
#pragma unroll
    for (int e = 0; e < 6; ++e) atomicAdd(dW + i * 6 + e, sW[e * D + i]);
    atomicAdd(db + i, sW[6 * D + i]);
  }
}
int launch_proj_in_bwd(const float* dx, const float* xt, float* dW, float* db, int B, int L, cudaStream_t s) {
  dim3 grid(ceil_div(L, TOKB), B);
  proj_in_bwd_kernel<<<grid, 256, 0, s>>>(dx, xt, dW, db, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// proj_audio backward helper: da_pre = da * silu'(pre) (bf16 operand for the wgrad GEMM), db = colsum.
__global__ void silu_bwd_kernel(const float* __restrict__ da, const float* __restrict__ pre,
                                __nv_bfloat16* __restrict__ dpre, float* __restrict__ db, int T) {
  __shared__ float sB[128];
  if (threadIdx.x < 128) sB[threadIdx.x] = 0.f;
  __syncthreads();
  const int c = threadIdx.x & 127, sub = threadIdx.x >> 7;  // 256 threads: 2 rows at a time
  float acc = 0.f;
  for (int t = blockIdx.x * 64 + sub; t < min(T, (int)(blockIdx.x + 1) * 64); t += 2) {
    const float z = pre[(size_t)t * 128 + c];
    const float sg = 1.0f / (1.0f + __expf(-z));
    const float d = da[(size_t)t * 128 + c] * (sg * (1.0f + z * (1.0f - sg)));
    dpre[(size_t)t * 128 + c] = __float2bfloat16(d);
    acc += d;
  }
  atomicAdd(&sB[c], acc);
  __syncthreads();
  if (threadIdx.x < 128) atomicAdd(db + threadIdx.x, sB[threadIdx.x]);
}
int launch_silu_bwd(const float* da, const float* pre, void* dpre, float* db, int T, cudaStream_t s) {
  silu_bwd_kernel<<<ceil_div(T, 64), 256, 0, s>>>(da, pre, static_cast<__nv_bfloat16*>(dpre), db, T);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// small dense layer backward (fp32): out = act(in W^T + b)
//   dpre = dout * act'(pre) ; dW[n][k] += sum_b dpre[b][n] in[b][k] ; db[n] += sum_b dpre ; din[b][k] += sum_n dpre W[n][k]
__global__ void linear_small_bwd_w_kernel(const float* __restrict__ dout, const float* __restrict__ out_act,
                                          const float* __restrict__ in, float* __restrict__ dW, float* __restrict__ db,
                                          float* __restrict__ dpre_out, int Bn, int N, int K, int silu) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float bsum = 0.f;
  for (int b = 0; b < Bn; ++b) {
    float d = dout[(size_t)b * N + n];
    if (silu) {
      // out_act holds the PRE-activation for silu layers
      const float z = out_act[(size_t)b * N + n];
      const float sg = 1.0f / (1.0f + __expf(-z));
      d *= sg * (1.0f + z * (1.0f - sg));
    }
    if (dpre_out != nullptr && lane == 0) dpre_out[(size_t)b * N + n] = d;
    bsum += d;
    for (int k = lane; k < K; k += 32) dW[(size_t)n * K + k] += d * in[(size_t)b * K + k];
  }
  if (lane == 0) db[n] += bsum;
}
__global__ void linear_small_bwd_in_kernel(const float* __restrict__ dpre, const float* __restrict__ W,
                                           float* __restrict__ din, int Bn, int N, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (k >= K) return;
  const int n0 = blockIdx.z * 64, n1 = min(N, n0 + 64);
  float acc = 0.f;
  for (int n = n0; n < n1; ++n) acc = fmaf(dpre[(size_t)b * N + n], __ldg(W + (size_t)n * K + k), acc);
  atomicAdd(din + (size_t)b * K + k, acc);
}
int launch_linear_small_bwd(const float* dout, const float* pre_or_null, const float* in, const float* W, float* dW,
                            float* db, float* din, float* dpre_scratch, int Bn, int N, int K, int silu,
                            cudaStream_t s) {
  linear_small_bwd_w_kernel<<<ceil_div(N, 8), 256, 0, s>>>(dout, pre_or_null, in, dW, db, silu ? dpre_scratch : nullptr,
                                                           Bn, N, K, silu);
  OSD_LAUNCHED();
  if (din != nullptr) {
    dim3 grid(ceil_div(K, 128), Bn, ceil_div(N, 64));
    linear_small_bwd_in_kernel<<<grid, 128, 0, s>>>(silu ? dpre_scratch : dout, W, din, Bn, N, K);
    OSD_LAUNCHED();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// inverse of pack_weight for gradients: padded fp32 [rows_dst(padded), cols_dst(padded)] -> parameter grad layout
__global__ void unpack_grad_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows_src, int cols_src,
                                   int rows_dst, int cols_dst, int split_at, int split_pad) {
  // dst is the PARAMETER gradient [rows_dst, cols_dst]; src the padded one [rows_src, cols_src]
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)rows_dst * cols_dst;
  if (i >= n) return;
  const int r = (int)(i / cols_dst), c = (int)(i % cols_dst);
  int sr = r;
  if (split_at > 0 && r >= split_at) sr = r - split_at + split_pad;
  dst[i] += src[(size_t)sr * cols_src + c];
}
int launch_unpack_grad(const float* src, float* dst, int rows_src, int cols_src, int rows_dst, int cols_dst,
                       int split_at, int split_pad, cudaStream_t s) {
  const size_t n = (size_t)rows_dst * cols_dst;
  unpack_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, rows_src, cols_src, rows_dst, cols_dst,
                                                                 split_at, split_pad);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
