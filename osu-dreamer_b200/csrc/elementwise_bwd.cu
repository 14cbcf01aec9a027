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
  static bool set = false;
  if (!set) {
    OSD_CUDA(cudaFuncSetAttribute(dwconv_prenorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    set = true;
  }
  dim3 grid(ceil_div(L, DWB_TOK * DWB_SUB), B);
  dwconv_prenorm_bwd_kernel<<<grid, 256, smem, s>>>(static_cast<const __nv_bfloat16*>(dz2),
                                                    static_cast<const __nv_bfloat16*>(hmod), x1, mod, wconv, dx, dmod,
                                                    dw, db, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SwiGLU + RMSNorm(1365) backward: dvg from dhn; dbvg (padded layout) = colsum(dvg).
static constexpr int HID = 1365, HIDP = 1408;
__global__ void __launch_bounds__(256) swiglu_norm_bwd_kernel(const __nv_bfloat16* __restrict__ vg,
                                                              const __nv_bfloat16* __restrict__ dhn,
                                                              const float* __restrict__ rinv,
                                                              __nv_bfloat16* __restrict__ dvg, float* __restrict__ dbvg,
                                                              int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * TOKB;
  for (int rr = warp; rr < TOKB; rr += 8) {
    const int t = t0 + rr;
    if (t >= T) break;
    const __nv_bfloat16* row = vg + (size_t)t * 2 * HIDP;
    const __nv_bfloat16* drow = dhn + (size_t)t * HIDP;
    const float r = rinv[t];
    // two passes over the row (the second one hits L1/L2): keeps the kernel at ~50 registers so that enough
    // warps are resident to cover HBM latency
    float dot = 0.f;
#pragma unroll 1
    for (int i = 0; i < 11; ++i) {
      const int c0 = i * 128 + lane * 4;
      const uint2 va = *reinterpret_cast<const uint2*>(row + c0);
      const uint2 ga = *reinterpret_cast<const uint2*>(row + HIDP + c0);
      const uint2 da = *reinterpret_cast<const uint2*>(drow + c0);
      const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&va);
      const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&ga);
      const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&da);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 v2 = __bfloat1622float2(ap[h]), g2 = __bfloat1622float2(bp[h]), d2 = __bfloat1622float2(dp[h]);
        dot = fmaf(d2.x, v2.x * silu_f(g2.x), dot);
        dot = fmaf(d2.y, v2.y * silu_f(g2.y), dot);
      }
    }
    dot = warp_sum(dot) * r * (1.0f / HID);  // mean(dhn * hn), hn = hs * r
#pragma unroll 1
    for (int i = 0; i < 11; ++i) {
      const int c0 = i * 128 + lane * 4;
      const uint2 va = *reinterpret_cast<const uint2*>(row + c0);
      const uint2 ga = *reinterpret_cast<const uint2*>(row + HIDP + c0);
      const uint2 da = *reinterpret_cast<const uint2*>(drow + c0);
      const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&va);
      const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&ga);
      const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&da);
      float ov[4], og[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 v2 = __bfloat1622float2(ap[h]), g2 = __bfloat1622float2(bp[h]), d2 = __bfloat1622float2(dp[h]);
        const float vv[2] = {v2.x, v2.y}, gg[2] = {g2.x, g2.y}, dd[2] = {d2.x, d2.y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float sg = 1.0f / (1.0f + __expf(-gg[e]));
          const float sil = gg[e] * sg;
          const float dhs = r * (dd[e] - vv[e] * sil * r * dot);
          ov[2 * h + e] = dhs * sil;
          og[2 * h + e] = dhs * vv[e] * (sg * (1.0f + gg[e] * (1.0f - sg)));
        }
      }
      *reinterpret_cast<uint2*>(dvg + (size_t)t * 2 * HIDP + c0) = make_uint2(pack_bf16(ov[0], ov[1]), pack_bf16(ov[2], ov[3]));
      *reinterpret_cast<uint2*>(dvg + (size_t)t * 2 * HIDP + HIDP + c0) =
          make_uint2(pack_bf16(og[0], og[1]), pack_bf16(og[2], og[3]));
    }
  }
}
// column sums of a bf16 [T, N] matrix (N % 8 == 0) accumulated into out[N] (fp32 atomics, one per column per block)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ m, float* __restrict__ out,
                                                          int T, int N) {
  const int t0 = blockIdx.x * 128;
  const int t1 = min(T, t0 + 128);
  for (int v = threadIdx.x; v < N / 8; v += 256) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = t0; t < t1; ++t) {
      const uint4 q = *reinterpret_cast<const uint4*>(m + (size_t)t * N + v * 8);
      const __nv_bfloat162* qp = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int k = 0; k < 4; ++k) a[2 * k] += __low2float(qp[k]), a[2 * k + 1] += __high2float(qp[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(out + v * 8 + k, a[k]);
  }
}
int launch_colsum_bf16(const void* m, float* out, int T, int N, cudaStream_t s) {
  OSD_CHECK(N % 8 == 0, "colsum: N must be a multiple of 8");
  colsum_bf16_kernel<<<ceil_div(T, 128), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(m), out, T, N);
  OSD_LAUNCHED();
  return 0;
}
int launch_swiglu_norm_bwd(const void* vg, const void* dhn, const float* rinv, void* dvg, float* dbvg, int T,
                           cudaStream_t s) {
  swiglu_norm_bwd_kernel<<<ceil_div(T, TOKB), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(vg),
                                                           static_cast<const __nv_bfloat16*>(dhn), rinv,
                                                           static_cast<__nv_bfloat16*>(dvg), dbvg, T);
  OSD_LAUNCHED();
  return launch_colsum_bf16(dvg, dbvg, T, 2 * HIDP, s);
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm(64)*w + RoPE backward, in place on dqkv [T,3072] (dq | dk | dv) -> gradients of the raw
// projections; dqn_w / dkn_w [64]; dbqkv [3072] = colsum of the result.  16 lanes per 64-wide head chunk:
// lane j of the half-warp holds elements (2j, 2j+1) and (2j+32, 2j+33) -- the rope partner pairs.
__global__ void __launch_bounds__(256) qknorm_rope_bwd_kernel(__nv_bfloat16* __restrict__ dqkv,
                                                              const __nv_bfloat16* __restrict__ raw,
                                                              const float* __restrict__ rope,
                                                              const float* __restrict__ qw, const float* __restrict__ kw,
                                                              float* __restrict__ dqw, float* __restrict__ dkw,
                                                              int L) {
  __shared__ float sW[2][64];
  if (threadIdx.x < 128) sW[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, l0 = blockIdx.x * TOKB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, j = lane & 15;
  const float eps = 1.1920929e-07f;
  // weights for this lane's 4 elements, q and k
  float wq[4], wkk[4];
  wq[0] = qw[2 * j], wq[1] = qw[2 * j + 1], wq[2] = qw[2 * j + 32], wq[3] = qw[2 * j + 33];
  wkk[0] = kw[2 * j], wkk[1] = kw[2 * j + 1], wkk[2] = kw[2 * j + 32], wkk[3] = kw[2 * j + 33];
  float adw[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) adw[a][e] = 0.f;

  for (int rr = warp; rr < TOKB; rr += 8) {
    const int l = l0 + rr;
    if (l >= L) break;
    const size_t t = (size_t)b * L + l;
    __nv_bfloat16* drow = dqkv + t * 3072;
    const __nv_bfloat16* xrow = raw + t * 3072;
    const float2 c01 = *reinterpret_cast<const float2*>(rope + (size_t)l * 64 + 2 * j);
    const float2 s01 = *reinterpret_cast<const float2*>(rope + (size_t)l * 64 + 32 + 2 * j);
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int chunk = it * 2 + half;  // 0..31: 16 q heads then 16 k heads
      const int col = chunk * 64;
      const int isk = chunk >= 16;
      const __nv_bfloat162 g1 = *reinterpret_cast<const __nv_bfloat162*>(drow + col + 2 * j);
      const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(drow + col + 32 + 2 * j);
      const __nv_bfloat162 x1 = *reinterpret_cast<const __nv_bfloat162*>(xrow + col + 2 * j);
      const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(xrow + col + 32 + 2 * j);
      const float gy[4] = {__low2float(g1), __high2float(g1), __low2float(g2), __high2float(g2)};
      const float xx[4] = {__low2float(x1), __high2float(x1), __low2float(x2), __high2float(x2)};
      // inverse rotation: da = g1*c + g2*s ; db = -g1*s + g2*c
      float da[4];
      da[0] = gy[0] * c01.x + gy[2] * s01.x;
      da[1] = gy[1] * c01.y + gy[3] * s01.y;
      da[2] = -gy[0] * s01.x + gy[2] * c01.x;
      da[3] = -gy[1] * s01.y + gy[3] * c01.y;
      float ss = xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2] + xx[3] * xx[3];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float r = rsqrtf(ss * (1.0f / 64.0f) + eps);
      float n[4], dn[4];
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        n[e] = xx[e] * r;
        const float w = isk ? wkk[e] : wq[e];
        dn[e] = da[e] * w;
        adw[isk][e] = fmaf(da[e], n[e], adw[isk][e]);
        dot = fmaf(dn[e], n[e], dot);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      dot *= (1.0f / 64.0f);
      float dxr[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dxr[e] = r * (dn[e] - n[e] * dot);
      }
      *reinterpret_cast<__nv_bfloat162*>(drow + col + 2 * j) = __floats2bfloat162_rn(dxr[0], dxr[1]);
      *reinterpret_cast<__nv_bfloat162*>(drow + col + 32 + 2 * j) = __floats2bfloat162_rn(dxr[2], dxr[3]);
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    atomicAdd(&sW[a][2 * j], adw[a][0]);
    atomicAdd(&sW[a][2 * j + 1], adw[a][1]);
    atomicAdd(&sW[a][2 * j + 32], adw[a][2]);
    atomicAdd(&sW[a][2 * j + 33], adw[a][3]);
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(dqw + threadIdx.x, sW[0][threadIdx.x]);
  else if (threadIdx.x < 128) atomicAdd(dkw + threadIdx.x - 64, sW[1][threadIdx.x - 64]);
}
int launch_qknorm_rope_bwd(void* dqkv, const void* raw, const float* rope, const float* qw, const float* kw,
                           float* dqw, float* dkw, float* dbias, int B, int L, cudaStream_t s) {
  dim3 grid(ceil_div(L, TOKB), B);
  qknorm_rope_bwd_kernel<<<grid, 256, 0, s>>>(static_cast<__nv_bfloat16*>(dqkv), static_cast<const __nv_bfloat16*>(raw),
                                              rope, qw, kw, dqw, dkw, L);
  OSD_LAUNCHED();
  return launch_colsum_bf16(dqkv, dbias, B * L, 3072, s);  // dbqkv = column sums of the raw-projection gradients
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
