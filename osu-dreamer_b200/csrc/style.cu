// Style model inference (reference: osu_dreamer/models/style/model.py:72-119): the [B, 32] sphere-tracing sampler that
// runs right before `diffusion.sample` inside LDM.sample (models/inference/model.py:48-49), and its forward.
// The model is tiny (5.97 M parameters, 8 FiLM-modulated MLP blocks of width 256 / 1024 on ONE vector per sample), so
// the work is latency, not throughput: one CTA per sample keeps the activation in shared memory, streams the fp32
// weights from L2 (warp-per-output-row mat-vecs, 16-byte loads) and runs ALL sampler steps inside one launch; the
// label conditioning and the 8 FiLM vectors depend only on the labels and are computed once per call.
//   launches per `sample`: conditioning, probe forward, eta (mean over the batch), step loop  = 4 (reference: ~1000).
// 512 threads per CTA: 16 warps x 2 rows x up to 8 16-byte loads per lane keep ~130 KB of weights in flight per SM.
#include "kernels.cuh"
#include "ptx.cuh"

#include <math.h>

namespace osd {

static constexpr int SD_ = 32, SH = 256, SE = 1024, SF = 128, SL = 5, SDEPTH = 8;
static constexpr float S_EPS = 1e-6f, S_EPS32 = 1.1920929e-07f;

struct StyleW {
  const float *cond_w, *cond_b, *null_labels, *rff_W, *rff_b, *in_w, *in_b, *out_g, *out_w, *out_b, *u_w, *u_b;
  const float *film_w[SDEPTH], *film_b[SDEPTH], *b0_w[SDEPTH], *b0_b[SDEPTH], *b3_w[SDEPTH], *b3_b[SDEPTH];
};
static StyleW style_weights(const float* const* P) {
  StyleW w;
  w.cond_w = P[0]; w.cond_b = P[1]; w.null_labels = P[2]; w.rff_W = P[3]; w.rff_b = P[4];
  w.in_w = P[5]; w.in_b = P[6]; w.out_g = P[7]; w.out_w = P[8]; w.out_b = P[9]; w.u_w = P[10]; w.u_b = P[11];
  for (int i = 0; i < SDEPTH; ++i) {
    w.film_w[i] = P[12 + 2 * i];
    w.film_b[i] = P[13 + 2 * i];
    w.b0_w[i] = P[28 + 4 * i];
    w.b0_b[i] = P[29 + 4 * i];
    w.b3_w[i] = P[30 + 4 * i];
    w.b3_b[i] = P[31 + 4 * i];
  }
  return w;
}

// y[n] = act(b[n] + sum_k W[n][k] x[k]); W row-major [N][K] in global memory (L2-resident), x in shared memory.
// NW warps, two output rows per warp and pass, all loads of a pass issued before the math: the weights are streamed
// once per forward by ONE CTA, so what matters is bytes in flight (N is even everywhere: 32, 256, 768, 1024).
template <int K, bool SILU, int NW>
__device__ __forceinline__ void matvec(const float* __restrict__ W, const float* __restrict__ b, const float* x, float* y,
                                       int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = 2 * warp; n < N; n += 2 * NW) {
    const float* r0 = W + (size_t)n * K;
    const float* r1 = r0 + K;
    float a0 = 0.f, a1 = 0.f;
    if constexpr (K % 128 == 0) {
      constexpr int C = K / 128;
      float4 w0[C], w1[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        w0[c] = __ldg(reinterpret_cast<const float4*>(r0 + c * 128 + lane * 4));
        w1[c] = __ldg(reinterpret_cast<const float4*>(r1 + c * 128 + lane * 4));
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(x + c * 128 + lane * 4);
        a0 = fmaf(w0[c].x, xv.x, a0), a0 = fmaf(w0[c].y, xv.y, a0), a0 = fmaf(w0[c].z, xv.z, a0), a0 = fmaf(w0[c].w, xv.w, a0);
        a1 = fmaf(w1[c].x, xv.x, a1), a1 = fmaf(w1[c].y, xv.y, a1), a1 = fmaf(w1[c].z, xv.z, a1), a1 = fmaf(w1[c].w, xv.w, a1);
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        a0 = fmaf(__ldg(r0 + k), x[k], a0);
        a1 = fmaf(__ldg(r1 + k), x[k], a1);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      float v0 = a0 + b[n], v1 = a1 + b[n + 1];
      if (SILU) v0 = v0 / (1.0f + expf(-v0)), v1 = v1 / (1.0f + expf(-v1));
      y[n] = v0;
      y[n + 1] = v1;
    }
  }
}
// sum of the values held by threads 0..255 (one per thread; the other threads pass 0); red = 8 floats of shared memory
__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();  // red may still be read from the previous use
  if ((threadIdx.x & 31) == 0 && threadIdx.x < 256) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// conditioning (model.py:72-79, fourier_features.py:15-16) and the 8 FiLM vectors (model.py:91): mod [B][8][768]
__global__ void __launch_bounds__(256) style_cond_kernel(StyleW w, const float* __restrict__ labels, float* __restrict__ mod) {
  __shared__ float ff[SL][SF];
  __shared__ __align__(16) float c[SH];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < SL * SF; i += 256) {
    const int n = i / SF, f = i % SF;
    const float lab = labels[b * SL + n];
    ff[n][f] = sqrtf(2.0f / SF) * cosf(lab / 10.0f * w.rff_W[f] + w.rff_b[f]);
  }
  __syncthreads();
  float acc = 0.f;
  for (int n = 0; n < SL; ++n) {
    if (labels[b * SL + n] < 0.f) {
      acc += w.null_labels[n * SH + tid];
    } else {
      float h = 0.f;
      const float* wn = w.cond_w + (size_t)n * SF * SH + tid;
#pragma unroll 8
      for (int f = 0; f < SF; ++f) h = fmaf(ff[n][f], wn[(size_t)f * SH], h);
      acc += h + w.cond_b[n * SH + tid];
    }
  }
  c[tid] = acc;
  __syncthreads();
  for (int i = 0; i < SDEPTH; ++i) {
    // straight to global memory (y pointer may be global for matvec: plain stores)
    matvec<SH, false, 8>(w.film_w[i], w.film_b[i], c, mod + ((size_t)b * SDEPTH + i) * 3 * SH, 3 * SH);
  }
}

// one forward of the trunk for the sample held by this CTA: s (shared, 32) -> u (returned), v (shared, 32).
// RUN_T threads: all warps stream weights in the mat-vecs, threads 0..255 own one channel each in the element-wise steps.
static constexpr int RUN_T = 512, RUN_W = RUN_T / 32;
__device__ float style_forward_dev(const StyleW& w, const float* mod, const float* s, float* v, float* x, float* h, float* e,
                                   float* red) {
  const int tid = threadIdx.x;
  const bool own = tid < SH;
  matvec<SD_, false, RUN_W>(w.in_w, w.in_b, s, x, SH);
  __syncthreads();
  for (int i = 0; i < SDEPTH; ++i) {
    const float* m = mod + (size_t)i * 3 * SH;
    const float xv = own ? x[tid] : 0.f;
    const float inv = rsqrtf(block_sum256(xv * xv, red) * (1.0f / SH) + S_EPS);
    if (own) h[tid] = xv * inv * (1.0f + m[tid]) + m[SH + tid];
    __syncthreads();
    matvec<SH, true, RUN_W>(w.b0_w[i], w.b0_b[i], h, e, SE);
    __syncthreads();
    matvec<SE, false, RUN_W>(w.b3_w[i], w.b3_b[i], e, h, SH);
    __syncthreads();
    const float hv = own ? h[tid] : 0.f;
    const float inv2 = rsqrtf(block_sum256(hv * hv, red) * (1.0f / SH) + S_EPS);
    if (own) x[tid] = xv + hv * inv2 * m[2 * SH + tid];
    __syncthreads();
  }
  const float xv = own ? x[tid] : 0.f;
  const float ms = block_sum256(xv * xv, red) * (1.0f / SH);
  if (own) h[tid] = xv * rsqrtf(ms + S_EPS32) * w.out_g[tid];  // nn.RMSNorm(h_dim), eps = finfo(fp32).eps (model.py:48)
  const float un = own ? xv * rsqrtf(ms + S_EPS) * w.u_w[tid] : 0.f;  // u_out(rms_norm(x)) (model.py:98)
  const float udot = block_sum256(un, red);  // (its barriers also publish h)
  matvec<SH, false, RUN_W>(w.out_w, w.out_b, h, v, SD_);
  __syncthreads();
  const float zv = udot + w.u_b[0];
  const float sp = zv > 20.f ? zv : log1pf(expf(zv));  // F.softplus (threshold 20)
  return sqrtf(2.0f * SD_) * sp;
}

// steps == 0: plain forward (u, v written out).  steps > 0: s <- s - eta * u * v, `steps` times, eta = eta_u0[0]
__global__ void __launch_bounds__(RUN_T, 1) style_run_kernel(StyleW w, const float* __restrict__ mod_all, float* __restrict__ s_io,
                                                          float* __restrict__ u_out, float* __restrict__ v_out,
                                                          const float* __restrict__ eta_u0, int steps) {
  __shared__ __align__(16) float s[SD_], v[SD_], x[SH], h[SH], e[SE], red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* mod = mod_all + (size_t)b * SDEPTH * 3 * SH;
  if (tid < SD_) s[tid] = s_io[b * SD_ + tid];
  __syncthreads();
  if (steps == 0) {
    const float u = style_forward_dev(w, mod, s, v, x, h, e, red);
    if (tid == 0) u_out[b] = u;
    if (v_out != nullptr && tid < SD_) v_out[b * SD_ + tid] = v[tid];
    return;
  }
  const float eta = eta_u0[0];
  for (int it = 0; it < steps; ++it) {
    const float u = style_forward_dev(w, mod, s, v, x, h, e, red);
    if (tid < SD_) s[tid] -= eta * u * v[tid];
    __syncthreads();
  }
  if (tid < SD_) s_io[b * SD_ + tid] = s[tid];
}

// eta = 1 - (sqrt(c0) / max(u0, sqrt(c0) + 1e-6))^(1/N), u0 = mean(u)  (model.py:110-111); out = {eta, u0}
__global__ void style_eta_kernel(const float* __restrict__ u, int B, float sqrt_c0, int num_steps, float* __restrict__ out) {
  float a = 0.f;
  for (int i = threadIdx.x; i < B; i += 32) a += u[i];
  a = warp_sum(a) / (float)B;
  if (threadIdx.x == 0) {
    out[0] = 1.0f - powf(sqrt_c0 / fmaxf(a, sqrt_c0 + 1e-6f), 1.0f / (float)num_steps);
    out[1] = a;
  }
}

size_t style_scratch_floats(int B) { return (size_t)B * SDEPTH * 3 * SH + (size_t)B + 8; }

int launch_style_forward(const float* const* params, const float* st, const float* labels, float* u, float* v, float* scratch,
                         int B, cudaStream_t s) {
  OSD_CHECK(params && st && labels && u && v && scratch && B > 0, "style_forward: bad arguments");
  const StyleW w = style_weights(params);
  style_cond_kernel<<<B, 256, 0, s>>>(w, labels, scratch);
  OSD_LAUNCHED();
  style_run_kernel<<<B, RUN_T, 0, s>>>(w, scratch, const_cast<float*>(st), u, v, nullptr, 0);
  OSD_LAUNCHED();
  return 0;
}

int launch_style_sample(const float* const* params, const float* labels, float* s_io, int num_steps, float* scratch,
                        float* eta_u0_out, int B, cudaStream_t s) {
  OSD_CHECK(params && labels && s_io && scratch && B > 0 && num_steps >= 1, "style_sample: bad arguments");
  const StyleW w = style_weights(params);
  float* mod = scratch;
  float* u = scratch + (size_t)B * SDEPTH * 3 * SH;
  float* eta = u + B;
  const float d0_sq = 2.0f * SD_;
  const float t99 = 1.0f / (1.0f + expf(-2.3263478740408408f));
  const float c0 = (1.0f - t99) * (1.0f - t99) * d0_sq;  // model.py:34-39
  style_cond_kernel<<<B, 256, 0, s>>>(w, labels, mod);
  OSD_LAUNCHED();
  style_run_kernel<<<B, RUN_T, 0, s>>>(w, mod, s_io, u, nullptr, nullptr, 0);
  OSD_LAUNCHED();
  style_eta_kernel<<<1, 32, 0, s>>>(u, B, sqrtf(c0), num_steps, eta);
  OSD_LAUNCHED();
  style_run_kernel<<<B, RUN_T, 0, s>>>(w, mod, s_io, u, nullptr, eta, num_steps);
  OSD_LAUNCHED();
  if (eta_u0_out != nullptr) OSD_CUDA(cudaMemcpyAsync(eta_u0_out, eta, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // namespace osd
