// Latent model, inference half (reference: osu_dreamer/models/latent/{model,unet,spec_features}.py): the
// `audio_encoder` that runs before `diffusion.sample` and the `decode` that runs after it inside LDM.sample
// (models/inference/model.py:47,51).  Width 128, fp32, channels-first [B, C, L] like the reference; ~0.7 TFLOP per song
// against 179 TFLOP per sampled latent, so these kernels are written for exact fp32 parity first (CUDA cores, fp32
// accumulation in the reference's order of operations up to summation order), not for the tensor pipe.
//   lat_block      one residual SwiGLU block of `layer` (unet.py:50-54): RMSNorm*gamma -> FiLM -> depthwise conv k=5 ->
//                  1x1 128->682 -> v*silu(g) -> RMSNorm(341) -> 1x1 341->128 -> RMSNorm*gamma -> *(1+gate) -> +x, fused per
//                  32-token tile in shared memory
//   lat_rmsnorm    RMSNorm over the channel dim with gamma (+ SiLU): out_norm, the stem norms, the mixer norm
//   lat_conv1x1    pointwise conv / linear with optional activation (mixers, proj_emb, proj_out, stem, FiLM, heads)
//   lat_conv2d     the two strided stem convolutions of SpecFeatures (kernel (kh,3), stride (sh,1), padding (1,1))
//   lat_down3 / lat_up3   depthwise conv k=3 + AvgPool(3) / nearest upsample x3 + depthwise conv k=3
//   lat_mix        x + p * g (mixer, unet.py:126)
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

static constexpr int LC = 128, LH = 341, LTL = 32, LHALO = 2, LTT = LTL + 2 * LHALO;
static constexpr float L_EPS = 1e-6f;

struct LatBlockW {
  const float *g1, *dw_w, *dw_b, *w1, *b1, *w2, *b2, *g2;
};

// dynamic smem layout (floats): xs [128][36] | hs [128][36] | zs [128][32] | vg [682][32] | part [8][36] | inv [36]
static constexpr int LB_XS = 0, LB_HS = LB_XS + LC * LTT, LB_ZS = LB_HS + LC * LTT, LB_VG = LB_ZS + LC * LTL,
                     LB_PART = LB_VG + 2 * LH * LTL, LB_INV = LB_PART + 8 * LTT, LB_FLOATS = LB_INV + LTT;

// per-token sum over a set of rows held by the warps -> inv[tt] = rsqrt(sum / n + eps); all 256 threads call it
__device__ __forceinline__ void lat_finish_stat(float* part, float* inv, float mine0, float mine1, float n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  part[warp * LTT + lane] = mine0;
  if (lane < LTT - 32) part[warp * LTT + 32 + lane] = mine1;
  __syncthreads();
  if (threadIdx.x < LTT) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w * LTT + threadIdx.x];
    inv[threadIdx.x] = rsqrtf(t / n + L_EPS);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) lat_block_kernel(const float* __restrict__ x, float* __restrict__ y, LatBlockW w,
                                                           const float* __restrict__ film /*[B][384] or null*/, int L) {
  extern __shared__ __align__(16) float sm[];
  float *xs = sm + LB_XS, *hs = sm + LB_HS, *zs = sm + LB_ZS, *vg = sm + LB_VG, *part = sm + LB_PART, *inv = sm + LB_INV;
  const int b = blockIdx.y, t0 = blockIdx.x * LTL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (size_t)b * LC * L;
  // P0: x tile with halo (zeros outside the sequence)
  for (int c = warp; c < LC; c += 8)
    for (int tt = lane; tt < LTT; tt += 32) {
      const int t = t0 - LHALO + tt;
      xs[c * LTT + tt] = (t >= 0 && t < L) ? xb[(size_t)c * L + t] : 0.f;
    }
  __syncthreads();
  // P1: rms over the 128 channels per token
  {
    float s0 = 0.f, s1 = 0.f;
    for (int c = warp; c < LC; c += 8) {
      const float a = xs[c * LTT + lane];
      s0 = fmaf(a, a, s0);
      if (lane < LTT - 32) {
        const float a1 = xs[c * LTT + 32 + lane];
        s1 = fmaf(a1, a1, s1);
      }
    }
    lat_finish_stat(part, inv, s0, s1, (float)LC);
  }
  // P2: h = norm(x) * gamma * (1 + scale) + shift, zero outside the sequence (the conv's zero padding)
  for (int c = warp; c < LC; c += 8) {
    const float g = w.g1[c];
    const float sc = film ? film[(size_t)b * 3 * LC + c] : 0.f, sh = film ? film[(size_t)b * 3 * LC + LC + c] : 0.f;
    for (int tt = lane; tt < LTT; tt += 32) {
      const int t = t0 - LHALO + tt;
      hs[c * LTT + tt] = (t >= 0 && t < L) ? (xs[c * LTT + tt] * inv[tt] * g) * (1.f + sc) + sh : 0.f;
    }
  }
  __syncthreads();
  // P3: depthwise conv k=5, pad 2
  for (int c = warp; c < LC; c += 8) {
    float a = w.dw_b[c];
#pragma unroll
    for (int k = 0; k < 5; ++k) a = fmaf(w.dw_w[c * 5 + k], hs[c * LTT + lane + k], a);
    zs[c * LTL + lane] = a;
  }
  __syncthreads();
  // P4: vg[n][t] = b1[n] + sum_c W1[n][c] z[c][t].  Register tile: a thread owns 4 tokens x 4 output rows, so every 4
  // channels cost 4 weight loads (16 B, L1-resident, shared by the 8 threads of a row group) + 4 activation loads (16 B from
  // shared memory, conflict-free) for 64 FMA -- 8 FMA per memory instruction (the first version: 1.3, LSU-bound at 19 TFLOP/s).
  // Each output is still one thread's sum over c in ascending order.
  {
    const int tq = threadIdx.x & 7, nr = threadIdx.x >> 3;  // tokens 4 tq .. 4 tq + 3, rows 4 (nr + 32 k) .. + 3
    for (int n0 = 4 * nr; n0 < 2 * LH; n0 += 128) {
      float acc[4][4];
      const float4* wr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = min(n0 + r, 2 * LH - 1);  // rows past 681 are computed on a clamped row and not stored
        wr[r] = reinterpret_cast<const float4*>(w.w1 + (size_t)n * LC);
        const float bb = w.b1[n];
        acc[r][0] = bb, acc[r][1] = bb, acc[r][2] = bb, acc[r][3] = bb;
      }
#pragma unroll 4
      for (int c4 = 0; c4 < LC / 4; ++c4) {
        float4 wv[4], zv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) wv[r] = __ldg(wr[r] + c4);
#pragma unroll
        for (int i = 0; i < 4; ++i) zv[i] = *reinterpret_cast<const float4*>(zs + (4 * c4 + i) * LTL + 4 * tq);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float ww[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[r][0] = fmaf(ww[i], zv[i].x, acc[r][0]);
            acc[r][1] = fmaf(ww[i], zv[i].y, acc[r][1]);
            acc[r][2] = fmaf(ww[i], zv[i].z, acc[r][2]);
            acc[r][3] = fmaf(ww[i], zv[i].w, acc[r][3]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (n0 + r < 2 * LH)
          *reinterpret_cast<float4*>(vg + (n0 + r) * LTL + 4 * tq) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
  }
  __syncthreads();
  // P5: h = v * silu(g) in place over the v rows; rms over the 341 hidden channels
  {
    float s0 = 0.f;
    for (int j = warp; j < LH; j += 8) {
      const float v = vg[j * LTL + lane], g = vg[(LH + j) * LTL + lane];
      const float h = v * (g / (1.0f + expf(-g)));
      vg[j * LTL + lane] = h;
      s0 = fmaf(h, h, s0);
    }
    lat_finish_stat(part, inv, s0, 0.f, (float)LH);
  }
  // P6: o[c][t] = b2[c] + inv[t] * sum_j W2[c][j] h[j][t]  -> zs.  Same register tile (4 tokens x 4 rows; the 128 rows are
  // one pass); rows of W2 are 341 floats, so the weights are scalar loads (4 per j) against one 16-byte activation load.
  {
    const int tq = threadIdx.x & 7, nr = threadIdx.x >> 3;
    const int c0 = 4 * nr;
    const float* wr0 = w.w2 + (size_t)c0 * LH;
    float acc[4][4] = {};
#pragma unroll 4
    for (int j = 0; j < LH; ++j) {
      const float4 hv = *reinterpret_cast<const float4*>(vg + j * LTL + 4 * tq);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float ww = __ldg(wr0 + (size_t)r * LH + j);
        acc[r][0] = fmaf(ww, hv.x, acc[r][0]);
        acc[r][1] = fmaf(ww, hv.y, acc[r][1]);
        acc[r][2] = fmaf(ww, hv.z, acc[r][2]);
        acc[r][3] = fmaf(ww, hv.w, acc[r][3]);
      }
    }
    const float4 iv = *reinterpret_cast<const float4*>(inv + 4 * tq);  // W2 (hn) = (W2 h) * inv: the norm is a per-token scalar
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float bb = w.b2[c0 + r];
      *reinterpret_cast<float4*>(zs + (c0 + r) * LTL + 4 * tq) =
          make_float4(fmaf(acc[r][0], iv.x, bb), fmaf(acc[r][1], iv.y, bb), fmaf(acc[r][2], iv.z, bb), fmaf(acc[r][3], iv.w, bb));
    }
  }
  __syncthreads();
  // P7: rms over channels of o, gain gamma, (1 + gate), residual
  {
    float s0 = 0.f;
    for (int c = warp; c < LC; c += 8) {
      const float a = zs[c * LTL + lane];
      s0 = fmaf(a, a, s0);
    }
    lat_finish_stat(part, inv, s0, 0.f, (float)LC);
  }
  const int t = t0 + lane;
  if (t < L) {
    float* yb = y + (size_t)b * LC * L;
    const float iv = inv[lane];
    for (int c = warp; c < LC; c += 8) {
      const float gt = film ? film[(size_t)b * 3 * LC + 2 * LC + c] : 0.f;
      yb[(size_t)c * L + t] = xs[c * LTT + LHALO + lane] + (zs[c * LTL + lane] * iv * w.g2[c]) * (1.f + gt);
    }
  }
}

int launch_lat_block(const float* x, float* y, const float* const* w8, const float* film, int B, int L, cudaStream_t s) {
  OSD_CHECK(x && y && w8 && x != y && B > 0 && L > 0, "lat_block: bad arguments (x and y must be distinct buffers)");
  LatBlockW w{w8[0], w8[1], w8[2], w8[3], w8[4], w8[5], w8[6], w8[7]};
  const int smem = LB_FLOATS * 4;
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(lat_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid(ceil_div(L, LTL), B);
  lat_block_kernel<<<grid, 256, smem, s>>>(x, y, w, film, L);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// y[b][c][n] = act(x[b][c][n] * rsqrt(mean_c x^2 + eps) * gamma[c]); N = product of the trailing dims; act 0 none, 1 SiLU
__global__ void lat_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, float* __restrict__ y, int C,
                                   long long N, long long total, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / N, n = i % N;
  const float* p = x + b * C * N + n;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = p[(long long)c * N];
    ss = fmaf(v, v, ss);
  }
  const float inv = rsqrtf(ss / (float)C + L_EPS);
  float* q = y + b * C * N + n;
  for (int c = 0; c < C; ++c) {
    float v = p[(long long)c * N] * inv;
    if (gamma != nullptr) v *= gamma[c];
    if (act == 1) v = v / (1.0f + expf(-v));
    q[(long long)c * N] = v;
  }
}
int launch_lat_rmsnorm(const float* x, const float* gamma, float* y, int B, int C, long long N, int act, cudaStream_t s) {
  OSD_CHECK(x && y && B > 0 && C > 0 && N > 0, "lat_rmsnorm: bad arguments");
  const long long total = (long long)B * N;
  lat_rmsnorm_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, gamma, y, C, N, total, act);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// y[b][o][n] = act(bias[o] + sum_i W[o][i] x[b][i][n]); act 0 none, 1 SiLU, 2 sigmoid on channels < act_channels
__global__ void __launch_bounds__(128) lat_conv1x1_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ y, int Cin,
                                                          int Cout, long long N, int act, int act_channels) {
  extern __shared__ float xs[];  // [Cin][128]
  const int b = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * 128, n = n0 + threadIdx.x;
  const float* xb = x + (long long)b * Cin * N;
  for (int i = 0; i < Cin; ++i) xs[i * 128 + threadIdx.x] = (n < N) ? xb[(long long)i * N + n] : 0.f;
  __syncthreads();
  if (n >= N) return;
  float* yb = y + (long long)b * Cout * N + n;
  for (int o = 0; o < Cout; ++o) {
    const float* wr = W + (long long)o * Cin;
    float a = bias ? bias[o] : 0.f;
    for (int i = 0; i < Cin; ++i) a = fmaf(__ldg(wr + i), xs[i * 128 + threadIdx.x], a);
    if (act == 1) a = a / (1.0f + expf(-a));
    else if (act == 2 && o < act_channels) a = 1.0f / (1.0f + expf(-a));
    yb[(long long)o * N] = a;
  }
}
// The same product when there are only a few positions (N < 32: the Linear layers -- FiLM projections [B,32] -> 384, label
// predictor): thread = one output (o, n) instead of one position, so a FiLM projection is 384 parallel 32-term sums and not one
// thread walking all of them (0.45 ms each, 24 of them in `decode`).  Same bias-first ascending fmaf chain per output.
__global__ void __launch_bounds__(128) lat_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ y, int Cin, int Cout,
                                                         int N, int act, int act_channels) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx >= Cout * N) return;
  const int o = idx / N, n = idx % N;
  const float* xb = x + (long long)b * Cin * N + n;
  const float* wr = W + (long long)o * Cin;
  float a = bias ? bias[o] : 0.f;
  for (int i = 0; i < Cin; ++i) a = fmaf(__ldg(wr + i), xb[(long long)i * N], a);
  if (act == 1) a = a / (1.0f + expf(-a));
  else if (act == 2 && o < act_channels) a = 1.0f / (1.0f + expf(-a));
  y[(long long)b * Cout * N + (long long)o * N + n] = a;
}

// The same product for the wide cases (Cin a multiple of 4, Cout >= 32: the decoder's mixers and gates at up to 160 k tokens),
// register-tiled like P4 of lat_block_kernel: block = 32 tokens, thread = 4 tokens x 4 output rows, 8 FMA per memory
// instruction instead of 0.5.  Every output is still bias + one thread's fmaf chain over i in ascending order, i.e. the result
// is bit-identical to the kernel above; outputs leave through shared memory so that the stores are 128-byte row segments.
__global__ void __launch_bounds__(256) lat_conv1x1_tiled_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                const float* __restrict__ bias, float* __restrict__ y, int Cin,
                                                                int Cout, long long N, int act, int act_channels) {
  extern __shared__ __align__(16) float xs[];  // [Cin][32], then reused as the output tile [128][33]
  const int b = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (long long)b * Cin * N;
  for (int i = warp; i < Cin; i += 8) xs[i * 32 + lane] = (n0 + lane < N) ? xb[(long long)i * N + n0 + lane] : 0.f;
  __syncthreads();
  const int tq = threadIdx.x & 7, nr = threadIdx.x >> 3;
  float* os = xs + Cin * 32;  // [128][33]
  for (int o0 = 0; o0 < Cout; o0 += 128) {
    float acc[4][4];
    const float4* wr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int o = min(o0 + 4 * nr + r, Cout - 1);  // rows past Cout - 1 are computed on a clamped row and not stored
      wr[r] = reinterpret_cast<const float4*>(W + (long long)o * Cin);
      const float bb = bias ? bias[o] : 0.f;
      acc[r][0] = bb, acc[r][1] = bb, acc[r][2] = bb, acc[r][3] = bb;
    }
#pragma unroll 4
    for (int c4 = 0; c4 < Cin / 4; ++c4) {
      float4 wv[4], zv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) wv[r] = __ldg(wr[r] + c4);
#pragma unroll
      for (int i = 0; i < 4; ++i) zv[i] = *reinterpret_cast<const float4*>(xs + (4 * c4 + i) * 32 + 4 * tq);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float ww[4] = {wv[r].x, wv[r].y, wv[r].z, wv[r].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[r][0] = fmaf(ww[i], zv[i].x, acc[r][0]);
          acc[r][1] = fmaf(ww[i], zv[i].y, acc[r][1]);
          acc[r][2] = fmaf(ww[i], zv[i].z, acc[r][2]);
          acc[r][3] = fmaf(ww[i], zv[i].w, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int o = o0 + 4 * nr + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = acc[r][j];
        if (act == 1) a = a / (1.0f + expf(-a));
        else if (act == 2 && o < act_channels) a = 1.0f / (1.0f + expf(-a));
        os[(4 * nr + r) * 33 + 4 * tq + j] = a;
      }
    }
    __syncthreads();
    if (n0 + lane < N) {
      float* yb = y + (long long)b * Cout * N + n0 + lane;
      for (int r = warp; r < 128 && o0 + r < Cout; r += 8) yb[(long long)(o0 + r) * N] = os[r * 33 + lane];
    }
    __syncthreads();
  }
}

int launch_lat_conv1x1(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, long long N,
                       int act, int act_channels, cudaStream_t s) {
  OSD_CHECK(x && W && y && B > 0 && Cin > 0 && Cin <= 256 && Cout > 0 && N > 0, "lat_conv1x1: bad arguments");
  if (N < 32) {
    dim3 grid_l((unsigned)ceil_div(Cout * (int)N, 128), B);
    lat_linear_kernel<<<grid_l, 128, 0, s>>>(x, W, bias, y, Cin, Cout, (int)N, act, act_channels);
    OSD_LAUNCHED();
    return 0;
  }
  if (Cin % 4 == 0 && Cout >= 32 && N >= 32 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    const int smem_t = (Cin * 32 + 128 * 33) * 4;
    static DeviceOnce once_t;  // per device; sized for the largest Cin accepted above
    if (once_t.first()) {
      OSD_CUDA(cudaFuncSetAttribute(lat_conv1x1_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (256 * 32 + 128 * 33) * 4));
    }
    dim3 grid_t((unsigned)((N + 31) / 32), B);
    lat_conv1x1_tiled_kernel<<<grid_t, 256, smem_t, s>>>(x, W, bias, y, Cin, Cout, N, act, act_channels);
    OSD_LAUNCHED();
    return 0;
  }
  const int smem = Cin * 128 * 4;
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(lat_conv1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 4));
  }
  dim3 grid((unsigned)((N + 127) / 128), B);
  lat_conv1x1_kernel<<<grid, 128, smem, s>>>(x, W, bias, y, Cin, Cout, N, act, act_channels);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// stem convolution: y[b][o][a][l] = bias[o] + sum_{i,p,q} W[o][i][p][q] x[b][i][a*sh + p - 1][l + q - 1], kernel (kh, 3),
// stride (sh, 1), padding (1, 1), zero padded; A_out = (A_in + 2 - kh) / sh + 1
__global__ void lat_conv2d_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                  float* __restrict__ y, int Cin, int Cout, int Ain, int Aout, int L, int kh, int sh,
                                  long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = (int)(idx % L);
  long long r = idx / L;
  const int a = (int)(r % Aout);
  r /= Aout;
  const int o = (int)(r % Cout);
  const int b = (int)(r / Cout);
  float acc = bias[o];
  for (int i = 0; i < Cin; ++i)
    for (int p = 0; p < kh; ++p) {
      const int ai = a * sh + p - 1;
      if (ai < 0 || ai >= Ain) continue;
      const float* xr = x + (((long long)b * Cin + i) * Ain + ai) * L;
      const float* wr = W + (((long long)o * Cin + i) * kh + p) * 3;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int li = l + q - 1;
        if (li >= 0 && li < L) acc = fmaf(wr[q], xr[li], acc);
      }
    }
  y[idx] = acc;
}
int launch_lat_conv2d(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, int Ain, int L,
                      int kh, int sh, cudaStream_t s) {
  OSD_CHECK(x && W && bias && y && B > 0 && Ain + 2 >= kh && sh > 0 && L > 0, "lat_conv2d: bad arguments");
  const int Aout = (Ain + 2 - kh) / sh + 1;
  const long long total = (long long)B * Cout * Aout * L;
  lat_conv2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, W, bias, y, Cin, Cout, Ain, Aout, L, kh, sh, total);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// down: depthwise conv k=3 pad 1, then AvgPool1d(3) (unet.py:60-64): y [B][C][L/3]
__global__ void lat_down3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ y, int C, int L, int Lo, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % Lo);
  const long long bc = idx / Lo;
  const int c = (int)(bc % C);
  const float* xr = x + bc * L;
  const float w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2], bb = bias[c];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int t = 3 * j + i;
    float v = bb;
    if (t - 1 >= 0) v = fmaf(w0, xr[t - 1], v);
    v = fmaf(w1, xr[t], v);
    if (t + 1 < L) v = fmaf(w2, xr[t + 1], v);
    acc += v;
  }
  y[idx] = acc / 3.0f;
}
// up: nearest upsample x3, then depthwise conv k=3 pad 1 (unet.py:81-85): y [B][C][3 l]
__global__ void lat_up3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ y, int C, int l, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Lo = 3 * l;
  const int t = (int)(idx % Lo);
  const long long bc = idx / Lo;
  const int c = (int)(bc % C);
  const float* xr = x + bc * l;
  float v = bias[c];
  if (t - 1 >= 0) v = fmaf(w[c * 3], xr[(t - 1) / 3], v);
  v = fmaf(w[c * 3 + 1], xr[t / 3], v);
  if (t + 1 < Lo) v = fmaf(w[c * 3 + 2], xr[(t + 1) / 3], v);
  y[idx] = v;
}
int launch_lat_down3(const float* x, const float* w, const float* bias, float* y, int B, int C, int L, cudaStream_t s) {
  OSD_CHECK(x && w && bias && y && B > 0 && C > 0 && L >= 3, "lat_down3: bad arguments");
  const int Lo = L / 3;
  const long long total = (long long)B * C * Lo;
  lat_down3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, w, bias, y, C, L, Lo, total);
  OSD_LAUNCHED();
  return 0;
}
int launch_lat_up3(const float* x, const float* w, const float* bias, float* y, int B, int C, int l, cudaStream_t s) {
  OSD_CHECK(x && w && bias && y && B > 0 && C > 0 && l > 0, "lat_up3: bad arguments");
  const long long total = (long long)B * C * 3 * l;
  lat_up3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, w, bias, y, C, l, total);
  OSD_LAUNCHED();
  return 0;
}

// y = x + p * g ; p may be broadcast over the batch (p_batch == 1: the audio skips of `predict`)
__global__ void lat_mix_kernel(const float* __restrict__ x, const float* __restrict__ p, const float* __restrict__ g,
                               float* __restrict__ y, long long per_sample, int p_bcast, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  y[i] = fmaf(p[p_bcast ? i % per_sample : i], g[i], x[i]);
}
int launch_lat_mix(const float* x, const float* p, const float* g, float* y, int B, long long per_sample, int p_batch,
                   cudaStream_t s) {
  OSD_CHECK(x && p && g && y && B > 0 && per_sample > 0 && (p_batch == 1 || p_batch == B), "lat_mix: bad arguments");
  const long long total = (long long)B * per_sample;
  lat_mix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, p, g, y, per_sample, p_batch == 1 && B > 1, total);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
