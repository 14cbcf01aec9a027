// Latent residual block on the tensor cores (reference: osu_dreamer/models/latent/unet.py:50-54, common/swiglu.py).
// The exact-fp32 kernel of latent.cu (CUDA cores, one fused 32-token tile) runs at 10 TFLOP/s and made `decode` +
// `audio_encoder` 78 % of `predict`; here the block's two 1x1 convolutions are tcgen05 GEMMs with split-TF32 operands
// (x = hi + lo with hi = tf32(x), products hi*hi + lo*hi + hi*lo accumulated in fp32: 3xTF32, ~1e-6 of the fp32 result --
// split-bf16 measured 1.5e-5 per layer and 7e-4 after the 48 blocks of encoder + decoder, too close to the 1e-3 bound) and
// the rest is three streaming kernels.  Tokens are rows ([B*L, C], channels-last) between the kernels, so the batch folds into M:
//   front : x [B,128,L] fp32 -> RMSNorm*g1 -> FiLM -> depthwise conv k=5 -> z (hi | lo) fp32 [T, 256]
//   GEMM1 : vg [T, 704] fp32 = z W1^T + b1           (W1 682 x 128 padded to 704 rows, (hi | lo) along K)
//   mid   : u = v * silu(g), hn = u * rsqrt(mean u^2 + eps) -> (hi | lo) fp32 [T, 768] (341 padded to 384)
//   GEMM2 : o [T, 128] fp32 = hn W2^T + b2
//   back  : y = x + RMSNorm(o) * g2 * (1 + gate), back to channels-first
#include "common.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

static constexpr int TC_C = 128, TC_H = 341, TC_HP = 384, TC_N1 = 704, TC_TL = 32, TC_HALO = 2, TC_TT = TC_TL + 2 * TC_HALO;
static constexpr float TC_EPS = 1e-6f;

// packed weights of one block: W1 (hi | lo) [704][256] | W2 (hi | lo) [128][768] | b1 padded [704], all fp32 storage
static constexpr size_t TC_W1_OFF = 0, TC_W2_OFF = TC_W1_OFF + (size_t)TC_N1 * 2 * TC_C * 4,
                        TC_B1_OFF = TC_W2_OFF + (size_t)TC_C * 2 * TC_HP * 4, TC_PACK_BYTES = TC_B1_OFF + (size_t)TC_N1 * 4;
size_t lat_tc_pack_bytes() { return TC_PACK_BYTES; }
// workspace: z [T][256] | vg [T][704] | hn [T][768] | o [T][128], fp32
size_t lat_tc_workspace_bytes(int B, int L) {
  const size_t T = (size_t)B * L;
  return T * (2 * TC_C + TC_N1 + 2 * TC_HP + TC_C) * 4 + 1024;
}

// hi = x rounded to tf32 (10 explicit mantissa bits, low 13 bits zero), lo = tf32(x - hi) (the difference is exact in fp32;
// rounding it here, to nearest, is half the error of the truncation the tensor core would apply to it)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}

__global__ void lat_tc_pack_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                   float* __restrict__ W1p, float* __restrict__ W2p, float* __restrict__ b1p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < TC_N1 * TC_C) {  // W1 [682][128] -> [704][hi 128 | lo 128]
    const int n = i / TC_C, c = i % TC_C;
    float hi, lo;
    split_tf32(n < 2 * TC_H ? w1[(size_t)n * TC_C + c] : 0.f, hi, lo);
    W1p[(size_t)n * 2 * TC_C + c] = hi;
    W1p[(size_t)n * 2 * TC_C + TC_C + c] = lo;
  }
  if (i < TC_C * TC_HP) {  // W2 [128][341] -> [128][hi 384 | lo 384]
    const int c = i / TC_HP, j = i % TC_HP;
    float hi, lo;
    split_tf32(j < TC_H ? w2[(size_t)c * TC_H + j] : 0.f, hi, lo);
    W2p[(size_t)c * 2 * TC_HP + j] = hi;
    W2p[(size_t)c * 2 * TC_HP + TC_HP + j] = lo;
  }
  if (i < TC_N1) b1p[i] = i < 2 * TC_H ? b1[i] : 0.f;
}

int launch_lat_tc_pack(const float* w1, const float* b1, const float* w2, void* packed, cudaStream_t s) {
  OSD_CHECK(w1 && b1 && w2 && packed, "lat_tc_pack: null argument");
  uint8_t* p = static_cast<uint8_t*>(packed);
  const int n = TC_N1 * TC_C;  // the largest of the three index spaces
  lat_tc_pack_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w1, b1, w2, reinterpret_cast<float*>(p + TC_W1_OFF),
                                                     reinterpret_cast<float*>(p + TC_W2_OFF),
                                                     reinterpret_cast<float*>(p + TC_B1_OFF));
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------ front
// One block = 32 tokens of one sample (+ 2 halo tokens each side), 256 threads; the arithmetic is P0-P3 of lat_block_kernel
// (same order of operations), the result leaves as (hi | lo) rows.
__global__ void __launch_bounds__(256) lat_tc_front_kernel(const float* __restrict__ x, const float* __restrict__ g1,
                                                           const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                           const float* __restrict__ film, float* __restrict__ z, int L) {
  __shared__ float xs[TC_C * TC_TT];        // [c][36]; becomes h in place
  __shared__ float zs[TC_C * (TC_TL + 1)];  // [c][33]
  __shared__ float part[8 * TC_TT], inv[TC_TT];
  const int b = blockIdx.y, t0 = blockIdx.x * TC_TL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (size_t)b * TC_C * L;
  for (int c = warp; c < TC_C; c += 8)
    for (int tt = lane; tt < TC_TT; tt += 32) {
      const int t = t0 - TC_HALO + tt;
      xs[c * TC_TT + tt] = (t >= 0 && t < L) ? xb[(size_t)c * L + t] : 0.f;
    }
  __syncthreads();
  {
    float s0 = 0.f, s1 = 0.f;
    for (int c = warp; c < TC_C; c += 8) {
      const float a = xs[c * TC_TT + lane];
      s0 = fmaf(a, a, s0);
      if (lane < TC_TT - 32) {
        const float a1 = xs[c * TC_TT + 32 + lane];
        s1 = fmaf(a1, a1, s1);
      }
    }
    part[warp * TC_TT + lane] = s0;
    if (lane < TC_TT - 32) part[warp * TC_TT + 32 + lane] = s1;
    __syncthreads();
    if (threadIdx.x < TC_TT) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w * TC_TT + threadIdx.x];
      inv[threadIdx.x] = rsqrtf(t / (float)TC_C + TC_EPS);
    }
    __syncthreads();
  }
  for (int c = warp; c < TC_C; c += 8) {
    const float g = g1[c];
    const float sc = film ? film[(size_t)b * 3 * TC_C + c] : 0.f, sh = film ? film[(size_t)b * 3 * TC_C + TC_C + c] : 0.f;
    for (int tt = lane; tt < TC_TT; tt += 32) {
      const int t = t0 - TC_HALO + tt;
      xs[c * TC_TT + tt] = (t >= 0 && t < L) ? (xs[c * TC_TT + tt] * inv[tt] * g) * (1.f + sc) + sh : 0.f;
    }
  }
  __syncthreads();
  for (int c = warp; c < TC_C; c += 8) {
    float a = dw_b[c];
#pragma unroll
    for (int k = 0; k < 5; ++k) a = fmaf(dw_w[c * 5 + k], xs[c * TC_TT + lane + k], a);
    zs[c * (TC_TL + 1) + lane] = a;
  }
  __syncthreads();
  // transpose out: warp w writes tokens w, w + 8, ...; lane = channel (4 x 32), conflict-free reads of the padded tile
  for (int tt = warp; tt < TC_TL; tt += 8) {
    const int t = t0 + tt;
    if (t >= L) break;
    float* row = z + ((size_t)b * L + t) * (2 * TC_C);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      float hi, lo;
      split_tf32(zs[c * (TC_TL + 1) + tt], hi, lo);
      row[c] = hi;
      row[TC_C + c] = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------ mid
// warp per token: u_j = v_j * silu(g_j) (j < 341; v = columns [0, 341), g = [341, 682) of vg), RMSNorm over the 341, (hi | lo)
__global__ void __launch_bounds__(256) lat_tc_mid_kernel(const float* __restrict__ vg, float* __restrict__ hn, long long T) {
  const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const float* row = vg + (size_t)t * TC_N1;
  float u[11];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int j = lane + 32 * k;
    float h = 0.f;
    if (j < TC_H) {
      const float v = row[j], g = row[TC_H + j];
      h = v * (g / (1.0f + expf(-g)));
    }
    u[k] = h;
    ss = fmaf(h, h, ss);
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)TC_H + TC_EPS);
  float* out = hn + (size_t)t * (2 * TC_HP);
#pragma unroll
  for (int k = 0; k < 12; ++k) {  // 384 = 12 x 32 columns; the padding columns [341, 384) are zeros
    const int j = lane + 32 * k;
    float hi, lo;
    split_tf32(k < 11 ? u[k < 11 ? k : 0] * inv : 0.f, hi, lo);
    out[j] = hi;
    out[TC_HP + j] = lo;
  }
}

// ------------------------------------------------------------------------------------------------ back
// block = 32 tokens of one sample: o rows in (coalesced), per-token RMS over the 128 channels, y out channels-first
__global__ void __launch_bounds__(256) lat_tc_back_kernel(const float* __restrict__ o, const float* __restrict__ x,
                                                          const float* __restrict__ g2, const float* __restrict__ film,
                                                          float* __restrict__ y, int L) {
  __shared__ float os[TC_TL * (TC_C + 1)];  // [t][129]
  __shared__ float inv[TC_TL];
  const int b = blockIdx.y, t0 = blockIdx.x * TC_TL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int tt = warp; tt < TC_TL; tt += 8) {
    const int t = t0 + tt;
    float ss = 0.f;
    if (t < L) {
      const float* row = o + ((size_t)b * L + t) * TC_C;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = row[lane + 32 * i];
        os[tt * (TC_C + 1) + lane + 32 * i] = a;
        ss = fmaf(a, a, ss);
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) inv[tt] = rsqrtf(ss / (float)TC_C + TC_EPS);
  }
  __syncthreads();
  const int t = t0 + lane;
  if (t < L) {
    const float iv = inv[lane];
    const float* xb = x + (size_t)b * TC_C * L;
    float* yb = y + (size_t)b * TC_C * L;
    for (int c = warp; c < TC_C; c += 8) {
      const float gt = film ? film[(size_t)b * 3 * TC_C + 2 * TC_C + c] : 0.f;
      yb[(size_t)c * L + t] = xb[(size_t)c * L + t] + (os[lane * (TC_C + 1) + c] * iv * g2[c]) * (1.f + gt);
    }
  }
}

// w8 as in launch_lat_block (the GEMM weights / b1 are read from `packed`, b2 from w8[6])
int launch_lat_block_tc(const float* x, float* y, const float* const* w8, const void* packed, const float* film, void* ws,
                        int B, int L, cudaStream_t s) {
  OSD_CHECK(x && y && w8 && packed && ws && x != y && B > 0 && L > 0, "lat_block_tc: bad arguments");
  const size_t T = (size_t)B * L;
  OSD_CHECK(T < (1ull << 31), "lat_block_tc: too many tokens");
  uint8_t* w = static_cast<uint8_t*>(ws);
  w += (1024 - (reinterpret_cast<uintptr_t>(w) & 1023)) & 1023;
  float* z = reinterpret_cast<float*>(w);
  float* vg = z + T * 2 * TC_C;
  float* hn = vg + T * TC_N1;
  float* o = hn + T * 2 * TC_HP;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  dim3 grid(ceil_div(L, TC_TL), B);
  lat_tc_front_kernel<<<grid, 256, 0, s>>>(x, w8[0], w8[1], w8[2], film, z, L);
  OSD_LAUNCHED();
  {
    GemmArgs g;
    g.A = z; g.lda = 2 * TC_C; g.B = pk + TC_W1_OFF; g.ldb = 2 * TC_C; g.M = (int)T; g.N = TC_N1; g.K = TC_C;
    g.elem = ELEM_TF32; g.split3 = 1; g.epi = EPI_STORE; g.C = vg; g.ldc = TC_N1; g.c_fp32 = 1;
    g.bias = reinterpret_cast<const float*>(pk + TC_B1_OFF);
    OSD_TRY(launch_gemm(g, s));
  }
  lat_tc_mid_kernel<<<(unsigned)ceil_div64((int64_t)T, 8), 256, 0, s>>>(vg, hn, (long long)T);
  OSD_LAUNCHED();
  {
    GemmArgs g;
    g.A = hn; g.lda = 2 * TC_HP; g.B = pk + TC_W2_OFF; g.ldb = 2 * TC_HP; g.M = (int)T; g.N = TC_C; g.K = TC_HP;
    g.elem = ELEM_TF32; g.split3 = 1; g.epi = EPI_STORE; g.C = o; g.ldc = TC_C; g.c_fp32 = 1; g.bias = w8[6];
    OSD_TRY(launch_gemm(g, s));
  }
  lat_tc_back_kernel<<<grid, 256, 0, s>>>(o, x, w8[7], film, y, L);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
