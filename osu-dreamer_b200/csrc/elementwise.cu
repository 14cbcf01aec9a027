// HBM-bound kernels of the denoiser forward path: layout changes, RMSNorm + adaLN modulation,
// depthwise conv, SwiGLU + RMSNorm(1365), gated residual, output heads.  Token-major activations
// [T = B*L, C] (C contiguous); one warp per token row unless noted.  All statistics in fp32.
#include "kernels.cuh"

#include <algorithm>
#include "ptx.cuh"

namespace osd {

static constexpr int D = 512;
static constexpr float RMS_EPS = 1e-6f;  // osu_dreamer/common/rms_norm.py:12

// 4 consecutive values -> bf16; in split mode the row is (hi | lo) with lo = bf16(x - float(hi)) at column +width
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* row, int width, int c0, const float (&o)[4], int split) {
  const uint32_t h0 = pack_bf16(o[0], o[1]), h1 = pack_bf16(o[2], o[3]);
  *reinterpret_cast<uint2*>(row + c0) = make_uint2(h0, h1);
  if (split) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&h0);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&h1);
    *reinterpret_cast<uint2*>(row + width + c0) =
        make_uint2(pack_bf16(o[0] - __low2float(a), o[1] - __high2float(a)),
                   pack_bf16(o[2] - __low2float(b), o[3] - __high2float(b)));
  }
}

// ------------------------------------------------------------------------------------------------
struct InvFreq {
  float v[32];
};
// rope[l][0][i] = cos(l*f_i), rope[l][1][i] = sin(l*f_i); the product l*f_i is rounded to fp32 first,
// as torch.outer(arange(N).float(), inv_freq) does in osu_dreamer/common/attn.py:16-21.
// A second copy follows at rope + L*64, transposed per block of 32 positions: ropeT[l/32][j][l%32][4] with j = 0..7 the
// cos float4 groups and j = 8..15 the sin groups.  The qkv GEMM epilogue (one thread per token row, 32 consecutive
// rows per warp) reads it with lane-consecutive 16-byte loads (4 wavefronts per request instead of 32).
__global__ void rope_table_kernel(InvFreq f, int L, int Lp, float* __restrict__ rope) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Lp * 32) return;
  const int l = idx >> 5, i = idx & 31;
  const float ang = __fmul_rn(static_cast<float>(l), f.v[i]);
  float s, c;
  sincosf(ang, &s, &c);
  if (l < L) {
    rope[(size_t)l * 64 + i] = c;
    rope[(size_t)l * 64 + 32 + i] = s;
  }
  float* t = rope + (size_t)L * 64 + (size_t)(l >> 5) * 2048;
  t[((i >> 2) * 32 + (l & 31)) * 4 + (i & 3)] = c;
  t[((8 + (i >> 2)) * 32 + (l & 31)) * 4 + (i & 3)] = s;
}
size_t rope_table_floats(int L) { return (size_t)L * 64 + (size_t)ceil_div(L, 32) * 2048; }
int launch_rope_table(const float* inv_freq_host, int L, float* rope, cudaStream_t stream) {
  OSD_CHECK(inv_freq_host && rope && L > 0, "rope_table: bad arguments");
  InvFreq f;
  for (int i = 0; i < 32; ++i) f.v[i] = inv_freq_host[i];
  const int Lp = ceil_div(L, 32) * 32;
  const int n = Lp * 32;
  rope_table_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(f, L, Lp, rope);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// [B, C, L] fp32 channels-first -> [B*L, C] token-major (bf16 or fp32).  32x32 smem transpose.
template <typename TOut>
__global__ void cf_to_tm_kernel(const float* __restrict__ in, TOut* __restrict__ out, int C, int L) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    tile[i][tx] = (c < C && l < L) ? in[((size_t)b * C + c) * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    if (l < L && c < C) out[((size_t)b * L + l) * C + c] = static_cast<TOut>(tile[tx][i]);
  }
}
__global__ void cf_to_tm_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int L) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    tile[i][tx] = (c < C && l < L) ? in[((size_t)b * C + c) * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    if (l < L && c < C) {
      const float v = tile[tx][i];
      const __nv_bfloat16 h = __float2bfloat16(v);
      out[((size_t)b * L + l) * 2 * C + c] = h;
      out[((size_t)b * L + l) * 2 * C + C + c] = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}
int launch_cf_to_tm_split(const float* in, void* out, int B, int C, int L, cudaStream_t stream) {
  dim3 grid(ceil_div(L, 32), ceil_div(C, 32), B), block(32, 8);
  cf_to_tm_split_kernel<<<grid, block, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), C, L);
  OSD_LAUNCHED();
  return 0;
}
int launch_cf_to_tm(const float* in, void* out, int out_bf16, int B, int C, int L, cudaStream_t stream) {
  dim3 grid(ceil_div(L, 32), ceil_div(C, 32), B), block(32, 8);
  if (out_bf16)
    cf_to_tm_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), C, L);
  else
    cf_to_tm_kernel<float><<<grid, block, 0, stream>>>(in, static_cast<float*>(out), C, L);
  OSD_LAUNCHED();
  return 0;
}

// [B*L, C] token-major (bf16 or fp32) -> [B, C, L] fp32 channels-first.
template <typename TIn>
__global__ void tm_to_cf_kernel(const TIn* __restrict__ in, float* __restrict__ out, int C, int L) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    tile[i][tx] = (c < C && l < L) ? static_cast<float>(in[((size_t)b * L + l) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    if (l < L && c < C) out[((size_t)b * C + c) * L + l] = tile[tx][i];
  }
}
int launch_tm_to_cf(const void* in, int in_fp32, float* out, int B, int C, int L, cudaStream_t stream) {
  dim3 grid(ceil_div(L, 32), ceil_div(C, 32), B), block(32, 8);
  if (in_fp32)
    tm_to_cf_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(in), out, C, L);
  else
    tm_to_cf_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, C, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Small dense layer on the conditioning vectors: out[b][n] = act(bias[n] + sum_k W[n][k] in[b][k]).
// fp32 throughout (proj_style, ssg1/ssg2, u_mod: model.py:46,66, backbone.py:62,66). One warp per n.
__global__ void linear_small_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ out, int Bn, int N, int K,
                                    int silu) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* w = W + (size_t)n * K;
  for (int b = 0; b < Bn; ++b) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(w + k), __ldg(in + (size_t)b * K + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + (bias ? bias[n] : 0.f);
      out[(size_t)b * N + n] = silu ? silu_f(v) : v;
    }
  }
}
int launch_linear_small(const float* in, const float* W, const float* bias, float* out, int Bn, int N, int K,
                        int silu, cudaStream_t stream) {
  linear_small_kernel<<<ceil_div(N, 8), 256, 0, stream>>>(in, W, bias, out, Bn, N, K, silu);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// proj_in (model.py:48,95): x[t][c] = b[c] + sum_e W[c][e] * xt[b][e][l];  xt channels-first fp32.
// A warp keeps the weights of its lanes' 16 channels in registers (16 x 6 + 16 bias) and walks 32 consecutive tokens: the six
// inputs of those tokens are six coalesced loads (lane j = token j) handed round by shuffles, so per token the warp issues
// 6 SHFL + 96 FMA + 4 16-byte stores -- the kernel is bound by its 2 KB/token of output (it was bound by 112 weight loads
// per token: 0.46 ms at B = 16, L = 8192 against 0.04 ms of HBM time).
static constexpr int PIN_TOK = 32;
__global__ void __launch_bounds__(256) proj_in_kernel(const float* __restrict__ xt, const float* __restrict__ W,
                                                      const float* __restrict__ bias, float* __restrict__ x, int L, int T) {
  const int lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * PIN_TOK;
  if (t0 >= T) return;
  float w[16][6], bb[16];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = v * 128 + lane * 4 + i;
      bb[4 * v + i] = __ldg(bias + c);
#pragma unroll
      for (int e = 0; e < 6; ++e) w[4 * v + i][e] = __ldg(W + c * 6 + e);
    }
  float in[6];
  {
    const int t = t0 + lane;
    const int b = t / L, l = t - b * L;  // a warp's 32 tokens may straddle two samples: per-lane addressing
#pragma unroll
    for (int e = 0; e < 6; ++e) in[e] = t < T ? __ldg(xt + ((size_t)b * 6 + e) * L + l) : 0.f;
  }
  const int n = min(PIN_TOK, T - t0);
  for (int k = 0; k < n; ++k) {
    float xin[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) xin[e] = __shfl_sync(0xffffffffu, in[e], k);
    float* row = x + (size_t)(t0 + k) * D;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float acc = bb[4 * v + i];
#pragma unroll
        for (int e = 0; e < 6; ++e) acc = fmaf(w[4 * v + i][e], xin[e], acc);
        o[i] = acc;
      }
      *reinterpret_cast<float4*>(row + v * 128 + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}
int launch_proj_in(const float* xt, const float* W, const float* bias, float* x, int B, int L, cudaStream_t stream) {
  const int T = B * L;
  proj_in_kernel<<<ceil_div(T, 8 * PIN_TOK), 256, 0, stream>>>(xt, W, bias, x, L, T);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// row helpers: a warp owns one 512-wide row; lane holds channels {v*128 + lane*4 + i}.
__device__ __forceinline__ void load_row_f32(const float* __restrict__ p, float (&r)[16], int lane) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 q = *reinterpret_cast<const float4*>(p + v * 128 + lane * 4);
    r[4 * v] = q.x, r[4 * v + 1] = q.y, r[4 * v + 2] = q.z, r[4 * v + 3] = q.w;
  }
}
__device__ __forceinline__ float row_inv_rms(const float (&r)[16]) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) ss = fmaf(r[i], r[i], ss);
  ss = warp_sum(ss);
  return rsqrtf(ss * (1.0f / D) + RMS_EPS);
}

// attn sub-block front end (backbone.py:76-78): z = rms_norm(x)*(1+scale)+shift + cl  -> bf16 / fp32.
// mod: [B, 1536] = (scale | shift | gate); cl: token-major [Tcl, 512] bf16, broadcast over batch when
// cl_bcast (the "#B = 1" audio of the predict path).
template <typename TOut>
__global__ void prenorm_mod_kernel(const float* __restrict__ x, const float* __restrict__ mod,
                                   const __nv_bfloat16* __restrict__ cl, TOut* __restrict__ z, int L, int T,
                                   int cl_bcast, int split, const float* __restrict__ cl_f32) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const int b = t / L;
  float r[16];
  load_row_f32(x + (size_t)t * D, r, lane);
  const float inv = row_inv_rms(r);
  const float* sc = mod + (size_t)b * 1536;
  const float* sh = sc + 512;
  const size_t tcl = cl_bcast ? (size_t)(t % L) : (size_t)t;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int c0 = v * 128 + lane * 4;
    const float4 s4 = *reinterpret_cast<const float4*>(sc + c0);
    const float4 h4 = *reinterpret_cast<const float4*>(sh + c0);
    float o[4] = {r[4 * v] * inv * (1.f + s4.x) + h4.x, r[4 * v + 1] * inv * (1.f + s4.y) + h4.y,
                  r[4 * v + 2] * inv * (1.f + s4.z) + h4.z, r[4 * v + 3] * inv * (1.f + s4.w) + h4.w};
    if (cl_f32 != nullptr) {
      const float4 c4 = *reinterpret_cast<const float4*>(cl_f32 + tcl * D + c0);
      o[0] += c4.x, o[1] += c4.y, o[2] += c4.z, o[3] += c4.w;
    } else if (cl != nullptr) {
      const uint2 c2 = *reinterpret_cast<const uint2*>(cl + tcl * D + c0);
      const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&c2.x);
      const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&c2.y);
      o[0] += __low2float(p0), o[1] += __high2float(p0), o[2] += __low2float(p1), o[3] += __high2float(p1);
    }
    if constexpr (sizeof(TOut) == 2) {
      store4_bf16(reinterpret_cast<__nv_bfloat16*>(z) + (size_t)t * (split ? 2 * D : D), D, c0, o, split);
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(z) + (size_t)t * D + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}
int launch_prenorm_mod(const float* x, const float* mod, const void* cl, void* z, int z_fp32, int B, int L,
                       int cl_bcast, cudaStream_t stream, int split, int cl_is_f32) {
  const int T = B * L;
  if (z_fp32)
    prenorm_mod_kernel<float><<<ceil_div(T, 8), 256, 0, stream>>>(x, mod, static_cast<const __nv_bfloat16*>(cl),
                                                                  static_cast<float*>(z), L, T, cl_bcast, 0, nullptr);
  else
    prenorm_mod_kernel<__nv_bfloat16><<<ceil_div(T, 8), 256, 0, stream>>>(
        x, mod, cl_is_f32 ? nullptr : static_cast<const __nv_bfloat16*>(cl), static_cast<__nv_bfloat16*>(z), L, T,
        cl_bcast, split, cl_is_f32 ? static_cast<const float*>(cl) : nullptr);
  OSD_LAUNCHED();
  return 0;
}

// gated residual (backbone.py:79-80, 85-86): x_out = x + rms_norm(h) * gate,  gate = mod[b][1024 + c].
__global__ void postnorm_gate_add_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                         const float* __restrict__ mod, float* __restrict__ x_out, int L, int T) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const int b = t / L;
  float r[16], xr[16];
  load_row_f32(h + (size_t)t * D, r, lane);
  load_row_f32(x + (size_t)t * D, xr, lane);
  const float inv = row_inv_rms(r);
  const float* gate = mod + (size_t)b * 1536 + 1024;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int c0 = v * 128 + lane * 4;
    const float4 g4 = *reinterpret_cast<const float4*>(gate + c0);
    *reinterpret_cast<float4*>(x_out + (size_t)t * D + c0) =
        make_float4(xr[4 * v] + r[4 * v] * inv * g4.x, xr[4 * v + 1] + r[4 * v + 1] * inv * g4.y,
                    xr[4 * v + 2] + r[4 * v + 2] * inv * g4.z, xr[4 * v + 3] + r[4 * v + 3] * inv * g4.w);
  }
}
int launch_postnorm_gate_add(const float* x, const float* h, const float* mod, float* x_out, int B, int L,
                             cudaStream_t stream) {
  const int T = B * L;
  postnorm_gate_add_kernel<<<ceil_div(T, 8), 256, 0, stream>>>(x, h, mod, x_out, L, T);
  OSD_LAUNCHED();
  return 0;
}

// ffn front end (backbone.py:83 + swiglu.py:20): hmod = rms_norm(x)*(1+scale)+shift, then the depthwise
// Conv1d(k=5, pad=2, groups=512) along the sequence (zero padding at the sample's ends) + bias.
// Block = 32 output tokens of one sample (+2 halo rows each side) x 512 channels, hmod staged in smem.
static constexpr int DW_TOK = 32;
template <typename TOut>
__global__ void __launch_bounds__(256) prenorm_mod_dwconv_kernel(
    const float* __restrict__ x, const float* __restrict__ mod, const float* __restrict__ wconv /*[512][5]*/,
    const float* __restrict__ bconv, TOut* __restrict__ z, __nv_bfloat16* __restrict__ hmod_out, int L, int split) {
  extern __shared__ __align__(128) float sh[];  // [DW_TOK + 4][512] + one mbarrier
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * DW_TOK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* sc = mod + (size_t)b * 1536;
  const float* shf = sc + 512;
  // the tile's rows (with 2 halo rows each side, clipped to the sample) are contiguous in x: ONE bulk copy brings the
  // 72 KB in, so a block has all of its bytes in flight from its first instruction (three blocks per SM overlap each
  // other's load / math / store phases); rows outside the sample are zero-filled
  uint64_t* bar = reinterpret_cast<uint64_t*>(sh + (DW_TOK + 4) * D);
  const int lo = max(l0 - 2, 0), hi = min(l0 + DW_TOK + 2, L);  // valid rows [lo, hi)
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    const uint32_t bytes = (uint32_t)(hi - lo) * D * 4;
    mbar_expect_tx(bar, bytes);
    bulk_load_1d(sh + (lo - (l0 - 2)) * D, x + ((size_t)b * L + lo) * D, bytes, bar);
  }
  for (int rr = warp; rr < DW_TOK + 4; rr += 8) {
    const int l = l0 + rr - 2;
    if (l < 0 || l >= L) {
      float* dst = sh + rr * D;
#pragma unroll
      for (int v = 0; v < 4; ++v) *reinterpret_cast<float4*>(dst + v * 128 + lane * 4) = make_float4(0, 0, 0, 0);
    }
  }
  __syncthreads();  // barrier initialised (and the zero rows written) before anyone waits / reads
  mbar_wait(bar, 0);
  for (int rr = warp; rr < DW_TOK + 4; rr += 8) {
    const int l = l0 + rr - 2;
    float* dst = sh + rr * D;
    if (l < 0 || l >= L) continue;
    const size_t t = (size_t)b * L + l;
    float r[16];
    load_row_f32(dst, r, lane);
    const float inv = row_inv_rms(r);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c0 = v * 128 + lane * 4;
      const float4 s4 = *reinterpret_cast<const float4*>(sc + c0);
      const float4 h4 = *reinterpret_cast<const float4*>(shf + c0);
      const float4 o = make_float4(r[4 * v] * inv * (1.f + s4.x) + h4.x, r[4 * v + 1] * inv * (1.f + s4.y) + h4.y,
                                   r[4 * v + 2] * inv * (1.f + s4.z) + h4.z, r[4 * v + 3] * inv * (1.f + s4.w) + h4.w);
      *reinterpret_cast<float4*>(dst + c0) = o;
      if (hmod_out != nullptr && rr >= 2 && rr < DW_TOK + 2)
        *reinterpret_cast<uint2*>(hmod_out + t * D + c0) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
  __syncthreads();
  // each thread: channels {threadIdx.x, threadIdx.x + 256}, all tokens
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = threadIdx.x + cc * 256;
    float w[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = __ldg(wconv + c * 5 + k);
    const float bb = __ldg(bconv + c);
    float win[5];
#pragma unroll
    for (int k = 0; k < 4; ++k) win[k + 1] = sh[k * D + c];
    for (int i = 0; i < DW_TOK; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) win[k] = win[k + 1];
      win[4] = sh[(i + 4) * D + c];
      const int l = l0 + i;
      if (l < L) {
        float acc = bb;
#pragma unroll
        for (int k = 0; k < 5; ++k) acc = fmaf(w[k], win[k], acc);
        if (split) {  // (hi | lo) row of width 2*D (bf16 only)
          const __nv_bfloat16 hh = __float2bfloat16(acc);
          __nv_bfloat16* zr = reinterpret_cast<__nv_bfloat16*>(z) + ((size_t)b * L + l) * 2 * D;
          zr[c] = hh;
          zr[D + c] = __float2bfloat16(acc - __bfloat162float(hh));
        } else {
          z[((size_t)b * L + l) * D + c] = static_cast<TOut>(acc);
        }
      }
    }
  }
}
int launch_prenorm_mod_dwconv(const float* x, const float* mod, const float* wconv, const float* bconv, void* z,
                              int z_fp32, void* hmod_out, int B, int L, cudaStream_t stream, int split) {
  dim3 grid(ceil_div(L, DW_TOK), B);
  const int smem = (DW_TOK + 4) * D * 4 + 16;
  if (z_fp32) {
    auto k = prenorm_mod_dwconv_kernel<float>;
    OSD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, 256, smem, stream>>>(x, mod, wconv, bconv, static_cast<float*>(z),
                                   static_cast<__nv_bfloat16*>(hmod_out), L, 0);
  } else {
    auto k = prenorm_mod_dwconv_kernel<__nv_bfloat16>;
    OSD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, 256, smem, stream>>>(x, mod, wconv, bconv, static_cast<__nv_bfloat16*>(z),
                                   static_cast<__nv_bfloat16*>(hmod_out), L, split);
  }
  OSD_LAUNCHED();
  return 0;
}

// SwiGLU + RMSNorm(1365, affine=False) (swiglu.py:28-30): vg = [v (1408 padded) | g (1408 padded)],
// hn = rms_norm_1365(v * silu(g)); padded columns are exactly zero (zero weight rows / bias).
static constexpr int HID = 1365, HIDP = 1408;
template <typename T>
__global__ void swiglu_norm_kernel(const T* __restrict__ vg, T* __restrict__ hn, float* __restrict__ rinv_out,
                                   int Tn, __nv_bfloat16* __restrict__ hn_split) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= Tn) return;
  const T* row = vg + (size_t)t * (2 * HIDP);
  float hs[44];  // 1408 / 32
  float ss = 0.f;
  if constexpr (sizeof(T) == 2) {
    // all 22 loads of the row are issued before the first use (5.6 KB in flight per warp): the kernel is a pure
    // HBM stream and was latency-bound with the loads interleaved into the math
    uint2 va[11], ga[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      va[i] = *reinterpret_cast<const uint2*>(row + i * 128 + lane * 4);
      ga[i] = *reinterpret_cast<const uint2*>(row + HIDP + i * 128 + lane * 4);
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      float v[4], g[4];
      unpack_bf16x4(va[i], v);
      unpack_bf16x4(ga[i], g);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = v[e] * g[e] * __fdividef(1.0f, 1.0f + __expf(-g[e]));  // silu with the fast reciprocal (2 ulp)
        hs[4 * i + e] = h;
        ss = fmaf(h, h, ss);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 11; ++i) {  // 11 chunks of 128 columns; lane owns 4 consecutive
      const int c0 = i * 128 + lane * 4;
      const float4 a = *reinterpret_cast<const float4*>(row + c0);
      const float4 bq = *reinterpret_cast<const float4*>(row + HIDP + c0);
      const float v[4] = {a.x, a.y, a.z, a.w}, g[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = v[e] * silu_f(g[e]);
        hs[4 * i + e] = h;
        ss = fmaf(h, h, ss);
      }
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss * (1.0f / HID) + RMS_EPS);
  if (rinv_out != nullptr && lane == 0) rinv_out[t] = inv;
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    const int c0 = i * 128 + lane * 4;
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint2*>(hn + (size_t)t * HIDP + c0) =
          make_uint2(pack_bf16(hs[4 * i] * inv, hs[4 * i + 1] * inv), pack_bf16(hs[4 * i + 2] * inv, hs[4 * i + 3] * inv));
    } else if (hn_split != nullptr) {
      const float o4[4] = {hs[4 * i] * inv, hs[4 * i + 1] * inv, hs[4 * i + 2] * inv, hs[4 * i + 3] * inv};
      store4_bf16(hn_split + (size_t)t * 2 * HIDP, HIDP, c0, o4, 1);
    } else {
      *reinterpret_cast<float4*>(hn + (size_t)t * HIDP + c0) =
          make_float4(hs[4 * i] * inv, hs[4 * i + 1] * inv, hs[4 * i + 2] * inv, hs[4 * i + 3] * inv);
    }
  }
}
// bf16 training / inference path: persistent, bulk-copy staged.  Each warp owns rows w, w + W, w + 2W ... and keeps
// SWS_ST of them in flight as 1-D bulk copies (cp.async.bulk, one 5632-byte vg row each) into its private
// shared-memory ring, so the bytes in flight (8 warps x 3 rows x 5.6 KB per SM) do not depend on registers or on how
// the compiler schedules loads; the math then reads the row from shared memory (conflict-free 8-byte accesses).
static constexpr int SWS_ST = 3, SWS_WARPS = 8, SWS_ROW = 2 * HIDP * 2;
__global__ void __launch_bounds__(SWS_WARPS * 32, 1) swiglu_norm_stream_kernel(const __nv_bfloat16* __restrict__ vg,
                                                                               __nv_bfloat16* __restrict__ hn,
                                                                               float* __restrict__ rinv_out, int Tn) {
  extern __shared__ __align__(128) uint8_t sws[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = sws + (size_t)warp * SWS_ST * SWS_ROW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sws + (size_t)SWS_WARPS * SWS_ST * SWS_ROW) + warp * SWS_ST;
  if (lane == 0) {
    for (int i = 0; i < SWS_ST; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  const int stride = gridDim.x * SWS_WARPS;
  const int first = blockIdx.x * SWS_WARPS + warp;
  auto issue = [&](int k) {  // k-th row of this warp -> ring stage k % SWS_ST (lane 0 only)
    const long long t = first + (long long)k * stride;
    if (t < Tn) {
      mbar_expect_tx(&bars[k % SWS_ST], SWS_ROW);
      bulk_load_1d(ring + (k % SWS_ST) * SWS_ROW, vg + (size_t)t * 2 * HIDP, SWS_ROW, &bars[k % SWS_ST]);
    }
  };
  if (lane == 0)
    for (int k = 0; k < SWS_ST - 1; ++k) issue(k);
  for (int k = 0;; ++k) {
    const long long t = first + (long long)k * stride;
    if (t >= Tn) break;
    if (lane == 0) {
      fence_proxy_async_smem();  // the stage refilled below was read (generic proxy) in the previous iteration
      issue(k + SWS_ST - 1);
    }
    mbar_wait(&bars[k % SWS_ST], (k / SWS_ST) & 1);
    const uint8_t* row = ring + (k % SWS_ST) * SWS_ROW;
    float hs[44];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      const uint2 va = *reinterpret_cast<const uint2*>(row + (i * 128 + lane * 4) * 2);
      const uint2 ga = *reinterpret_cast<const uint2*>(row + (HIDP + i * 128 + lane * 4) * 2);
      float v[4], g[4];
      unpack_bf16x4(va, v);
      unpack_bf16x4(ga, g);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = v[e] * g[e] * __fdividef(1.0f, 1.0f + __expf(-g[e]));  // silu with the fast reciprocal (2 ulp)
        hs[4 * i + e] = h;
        ss = fmaf(h, h, ss);
      }
    }
    __syncwarp();  // every lane has finished reading this stage before lane 0 may refill it
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss * (1.0f / HID) + RMS_EPS);
    if (rinv_out != nullptr && lane == 0) rinv_out[t] = inv;
#pragma unroll
    for (int i = 0; i < 11; ++i)
      *reinterpret_cast<uint2*>(hn + (size_t)t * HIDP + i * 128 + lane * 4) =
          make_uint2(pack_bf16(hs[4 * i] * inv, hs[4 * i + 1] * inv), pack_bf16(hs[4 * i + 2] * inv, hs[4 * i + 3] * inv));
  }
}
// is_fp32 = 2: fp32 vg in, bf16 (hi | lo) hn out [T, 2*HIDP]
int launch_swiglu_norm(const void* vg, void* hn, float* rinv_out, int is_fp32, int T, cudaStream_t stream) {
  if (is_fp32 == 2)
    swiglu_norm_kernel<float><<<ceil_div(T, 8), 256, 0, stream>>>(static_cast<const float*>(vg), nullptr, rinv_out, T,
                                                                  static_cast<__nv_bfloat16*>(hn));
  else if (is_fp32)
    swiglu_norm_kernel<float><<<ceil_div(T, 8), 256, 0, stream>>>(static_cast<const float*>(vg),
                                                                  static_cast<float*>(hn), rinv_out, T, nullptr);
  else {
    constexpr int smem = SWS_WARPS * SWS_ST * SWS_ROW + SWS_WARPS * SWS_ST * 8;
    static DeviceOnce once;
    if (once.first()) {
      OSD_CUDA(cudaFuncSetAttribute(swiglu_norm_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int grid = std::min(ceil_div(T, SWS_WARPS), num_sms());
    swiglu_norm_stream_kernel<<<grid, SWS_WARPS * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(vg),
                                                                     static_cast<__nv_bfloat16*>(hn), rinv_out, T);
  }
  OSD_LAUNCHED();
  return 0;
}

// final rms_norm (backbone.py:50) + proj_out (model.py:97): v[b][e][l] = bo[e] + sum_c Wo[e][c] * n[t][c],
// written channels-first fp32 [B, 6, L].
__global__ void final_norm_proj_out_kernel(const float* __restrict__ x, const float* __restrict__ Wo /*[6][512]*/,
                                           const float* __restrict__ bo, float* __restrict__ v, int L, int T) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  float r[16];
  load_row_f32(x + (size_t)t * D, r, lane);
  const float inv = row_inv_rms(r);
  float acc[6];
#pragma unroll
  for (int e = 0; e < 6; ++e) {
    float a = 0.f;
#pragma unroll
    for (int vv = 0; vv < 4; ++vv) {
      const float4 w = *reinterpret_cast<const float4*>(Wo + e * D + vv * 128 + lane * 4);
      a = fmaf(w.x, r[4 * vv], a);
      a = fmaf(w.y, r[4 * vv + 1], a);
      a = fmaf(w.z, r[4 * vv + 2], a);
      a = fmaf(w.w, r[4 * vv + 3], a);
    }
    acc[e] = warp_sum(a) * inv;
  }
  if (lane < 6) {
    const int b = t / L, l = t % L;
    float o = acc[0];
#pragma unroll
    for (int e = 1; e < 6; ++e) o = (lane == e) ? acc[e] : o;
    v[((size_t)b * 6 + lane) * L + l] = o + __ldg(bo + lane);
  }
}
int launch_final_norm_proj_out(const float* x, const float* Wo, const float* bo, float* v, int B, int L,
                               cudaStream_t stream) {
  const int T = B * L;
  final_norm_proj_out_kernel<<<ceil_div(T, 8), 256, 0, stream>>>(x, Wo, bo, v, L, T);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// u head (model.py:58-65,99): dwconv3(6) -> 1x1 6->64 -> SiLU -> dwconv3(64) -> 1x1 64->64 -> SiLU -> sum over l.
// Block = UH_TOK tokens of one sample (+1 halo each side) -> one partial row fpart[b][block][64]; u_final_kernel adds
// the rows in block order (no float atomics: the sampler is then bit-reproducible run to run -- a last-bit change of u
// is amplified by the bf16 roundings of the following forwards to ~1e-3 in individual outputs).
struct UHeadW {
  const float *w0, *b0;  // [6][3], [6]
  const float *w1, *b1;  // [64][6], [64]
  const float *w3, *b3;  // [64][3], [64]
  const float *w4, *b4;  // [64][64], [64]
};
static constexpr int UH_TOK = 32;
__global__ void __launch_bounds__(256) u_head_kernel(const float* __restrict__ xt, UHeadW w, float* __restrict__ fpart,
                                                     float* __restrict__ h1_save, float* __restrict__ h2pre_save, int L) {
  __shared__ float c1[UH_TOK + 2][6];
  __shared__ float h1[UH_TOK + 2][65];
  __shared__ float c2[UH_TOK][65];
  __shared__ float w4s[64][65];
  __shared__ float red[4][64];
  const int b = blockIdx.y, l0 = blockIdx.x * UH_TOK;
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * 64; i += 256) w4s[i >> 6][i & 63] = w.w4[i];
  // c1 = dwconv3 over l (zero padded) for positions l0-1 .. l0+UH_TOK
  for (int i = tid; i < (UH_TOK + 2) * 6; i += 256) {
    const int rr = i / 6, e = i % 6;
    const int l = l0 + rr - 1;
    float acc = 0.f;
    if (l >= 0 && l < L) {
      acc = w.b0[e];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ll = l + k - 1;
        if (ll >= 0 && ll < L) acc = fmaf(w.w0[e * 3 + k], xt[((size_t)b * 6 + e) * L + ll], acc);
      }
    }
    c1[rr][e] = acc;
  }
  __syncthreads();
  // h1 = silu(W1 c1 + b1); zero outside the sequence (it is the zero padding of the second dwconv)
  for (int i = tid; i < (UH_TOK + 2) * 64; i += 256) {
    const int rr = i >> 6, u = i & 63;
    const int l = l0 + rr - 1;
    float v = 0.f;
    if (l >= 0 && l < L) {
      float acc = w.b1[u];
#pragma unroll
      for (int e = 0; e < 6; ++e) acc = fmaf(w.w1[u * 6 + e], c1[rr][e], acc);
      v = silu_f(acc);
      if (h1_save != nullptr && rr >= 1 && rr <= UH_TOK) h1_save[((size_t)b * L + l) * 64 + u] = v;
    }
    h1[rr][u] = v;
  }
  __syncthreads();
  for (int i = tid; i < UH_TOK * 64; i += 256) {
    const int rr = i >> 6, u = i & 63;
    c2[rr][u] = w.b3[u] + w.w3[u * 3] * h1[rr][u] + w.w3[u * 3 + 1] * h1[rr + 1][u] + w.w3[u * 3 + 2] * h1[rr + 2][u];
  }
  __syncthreads();
  // h2 = silu(W4 c2 + b4), summed over this block's valid tokens. thread -> (u = tid & 63, token group = tid >> 6)
  {
    const int u = tid & 63, grp = tid >> 6;
    float part = 0.f;
    for (int rr = grp; rr < UH_TOK; rr += 4) {
      const int l = l0 + rr;
      if (l >= L) break;
      float acc = w.b4[u];
#pragma unroll 16
      for (int k = 0; k < 64; ++k) acc = fmaf(w4s[u][k], c2[rr][k], acc);
      if (h2pre_save != nullptr) h2pre_save[((size_t)b * L + l) * 64 + u] = acc;
      part += silu_f(acc);
    }
    red[grp][u] = part;
  }
  __syncthreads();
  if (tid < 64)
    fpart[((size_t)b * gridDim.x + blockIdx.x) * 64 + tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}
size_t u_head_partial_floats(int B, int L) { return (size_t)B * ceil_div(L, UH_TOK) * 64; }
int launch_u_head(const float* xt, const float* const* w8, float* fpart, float* h1_save, float* h2pre_save, int B,
                  int L, cudaStream_t stream) {
  UHeadW w{w8[0], w8[1], w8[2], w8[3], w8[4], w8[5], w8[6], w8[7]};
  dim3 grid(ceil_div(L, UH_TOK), B);
  u_head_kernel<<<grid, 256, 0, stream>>>(xt, w, fpart, h1_save, h2pre_save, L);
  OSD_LAUNCHED();
  return 0;
}

// u = u_scale * softplus(u_out(f * (1 + scale) + shift)),  f = fsum / L,  (scale | shift) = u_mod(cg)  (model.py:99-102)
// fsum[b][64] = sum over the u_head blocks of fpart[b][block][64], in block order (kept for the backward).
__global__ void u_final_kernel(const float* __restrict__ fpart, int nblk, float* __restrict__ fsum,
                               const float* __restrict__ umod /*[B][128]*/,
                               const float* __restrict__ wout /*[64]*/, const float* __restrict__ bout, float u_scale,
                               float inv_L, float* __restrict__ u, int Bn) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= Bn) return;
  float acc = 0.f;
  for (int i = lane; i < 64; i += 32) {
    float tot = 0.f;
    for (int k = 0; k < nblk; ++k) tot += fpart[((size_t)b * nblk + k) * 64 + i];
    fsum[(size_t)b * 64 + i] = tot;
    const float f = tot * inv_L;
    const float fm = f * (1.f + umod[(size_t)b * 128 + i]) + umod[(size_t)b * 128 + 64 + i];
    acc = fmaf(wout[i], fm, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float zv = acc + bout[0];
    const float sp = zv > 20.f ? zv : log1pf(expf(zv));  // F.softplus (threshold 20)
    u[b] = u_scale * sp;
  }
}
int launch_u_final(const float* fpart, float* fsum, const float* umod, const float* wout, const float* bout,
                   float u_scale, int L, float* u, int B, cudaStream_t stream) {
  u_final_kernel<<<ceil_div(B, 4), 128, 0, stream>>>(fpart, ceil_div(L, UH_TOK), fsum, umod, wout, bout, u_scale,
                                                     1.0f / (float)L, u, B);
  OSD_LAUNCHED();
  return 0;
}

// sampler update (model.py:136): x <- x - eta * u[b] * v   (channels-first [B, 6, L])
__global__ void sample_update_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ u,
                                     const float* __restrict__ eta, int per_sample, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / per_sample);
  x[i] = x[i] - eta[0] * u[b] * v[i];
}
int launch_sample_update(float* x, const float* v, const float* u, const float* eta_dev, int B, int L,
                         cudaStream_t stream) {
  const size_t n = (size_t)B * 6 * L;
  sample_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, v, u, eta_dev, 6 * L, n);
  OSD_LAUNCHED();
  return 0;
}

// eta = 1 - (sqrt(c0) / max(mean(u), sqrt(c0) + 1e-6)) ** (1/num_steps)   (model.py:131-132), on device.
__global__ void sample_eta_kernel(const float* __restrict__ u, int Bn, float sqrt_c0, float inv_steps,
                                  float* __restrict__ eta_out /*[2]: eta, u0*/) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < Bn; ++b) s += (double)u[b];
    const double u0 = s / Bn;
    const double den = fmax(u0, (double)sqrt_c0 + 1e-6);
    eta_out[0] = (float)(1.0 - pow((double)sqrt_c0 / den, (double)inv_steps));
    eta_out[1] = (float)u0;
  }
}
int launch_sample_eta(const float* u, int B, float sqrt_c0, int num_steps, float* eta_out, cudaStream_t stream) {
  sample_eta_kernel<<<1, 32, 0, stream>>>(u, B, sqrt_c0, 1.0f / (float)num_steps, eta_out);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Upper bound of the scaled attention scores in log2 units for one layer: q and k leave nn.RMSNorm(64) with
// ||.||_2 <= 8 before the elementwise gain (attn.py:71-72,77-78) and RoPE is a rotation, so
// |q.k| / sqrt(64) <= 8 max|w_q| max|w_k|.  Written as +inf when too large for a fixed-max softmax.
__global__ void qk_bound_kernel(const float* __restrict__ qw, const float* __restrict__ kw, float* __restrict__ out) {
  float a = fmaxf(fabsf(qw[threadIdx.x]), fabsf(qw[threadIdx.x + 32]));
  float b = fmaxf(fabsf(kw[threadIdx.x]), fabsf(kw[threadIdx.x + 32]));
  a = warp_max(a);
  b = warp_max(b);
  if (threadIdx.x == 0) {
    const float bound = 8.0f * a * b * 1.4426950408889634f;
    out[0] = (bound < 48.0f) ? bound : INFINITY;
  }
}
int launch_qk_bound(const float* qw, const float* kw, float* out, cudaStream_t stream) {
  qk_bound_kernel<<<1, 32, 0, stream>>>(qw, kw, out);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 parameter [rows_src, cols_src] -> operand dtype [rows_dst, cols_dst] with a row
// map (dst row r takes src row r - row_shift when inside [row_lo, row_hi)) and zero padding elsewhere.
template <typename TOut>
__global__ void pack_weight_kernel(const float* __restrict__ src, TOut* __restrict__ dst, int rows_src, int cols_src,
                                   int rows_dst, int cols_dst, int split_at, int split_pad) {
  // dst row r: if split_at > 0 rows [0, split_at) <- src [0, split_at); rows [split_pad, split_pad + split_at) <-
  // src [split_at, 2*split_at); everything else zero.  split_at == 0: plain copy of [rows_src, cols_src].
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)rows_dst * cols_dst;
  if (i >= n) return;
  const int r = (int)(i / cols_dst), c = (int)(i % cols_dst);
  int sr = -1;
  if (split_at > 0) {
    if (r < split_at)
      sr = r;
    else if (r >= split_pad && r < split_pad + split_at)
      sr = r - split_pad + split_at;
  } else if (r < rows_src) {
    sr = r;
  }
  float v = 0.f;
  if (sr >= 0 && c < cols_src) v = src[(size_t)sr * cols_src + c];
  dst[i] = static_cast<TOut>(v);
}
__global__ void pack_weight_split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows_src,
                                         int cols_src, int rows_dst, int cols_dst, int split_at, int split_pad) {
  // like pack_weight_kernel but every destination row is (hi | lo): [rows_dst, 2 * cols_dst]
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)rows_dst * cols_dst;
  if (i >= n) return;
  const int r = (int)(i / cols_dst), c = (int)(i % cols_dst);
  int sr = -1;
  if (split_at > 0) {
    if (r < split_at)
      sr = r;
    else if (r >= split_pad && r < split_pad + split_at)
      sr = r - split_pad + split_at;
  } else if (r < rows_src) {
    sr = r;
  }
  float v = 0.f;
  if (sr >= 0 && c < cols_src) v = src[(size_t)sr * cols_src + c];
  const __nv_bfloat16 h = __float2bfloat16(v);
  dst[(size_t)r * 2 * cols_dst + c] = h;
  dst[(size_t)r * 2 * cols_dst + cols_dst + c] = __float2bfloat16(v - __bfloat162float(h));
}
int launch_pack_weight_split(const float* src, void* dst, int rows_src, int cols_src, int rows_dst, int cols_dst,
                             int split_at, int split_pad, cudaStream_t stream) {
  const size_t n = (size_t)rows_dst * cols_dst;
  pack_weight_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
      src, static_cast<__nv_bfloat16*>(dst), rows_src, cols_src, rows_dst, cols_dst, split_at, split_pad);
  OSD_LAUNCHED();
  return 0;
}
int launch_pack_weight(const float* src, void* dst, int dst_fp32, int rows_src, int cols_src, int rows_dst,
                       int cols_dst, int split_at, int split_pad, cudaStream_t stream) {
  const size_t n = (size_t)rows_dst * cols_dst;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dst_fp32)
    pack_weight_kernel<float><<<grid, 256, 0, stream>>>(src, static_cast<float*>(dst), rows_src, cols_src, rows_dst,
                                                        cols_dst, split_at, split_pad);
  else
    pack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), rows_src,
                                                                cols_src, rows_dst, cols_dst, split_at, split_pad);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
