// Fused optimizer tail of fit-denoiser: global-norm gradient clipping
// (Lightning gradient_clip_val 1.0, model.yml:39) + torch.optim.AdamW step (train.py:111) + EMA update
// (swa_utils.get_ema_multi_avg_fn(.99), train.py:67,126) in one pass over flat fp32 buffers.
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ acc) {
  __shared__ float red[8];
  float s = 0.f;
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s = fmaf(v.x, v.x, s), s = fmaf(v.y, v.y, s), s = fmaf(v.z, v.z, s), s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(acc, (double)t);
  }
}
// scal[0] = total grad norm (after grad_scale), scal[1] = multiplier applied to raw gradients inside AdamW
__global__ void clip_coef_kernel(const double* __restrict__ acc, float grad_scale, float max_norm,
                                 float* __restrict__ scal) {
  const float norm = sqrtf((float)acc[0]) * grad_scale;
  float coef = 1.0f;
  if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (norm + 1e-6f));  // torch clip_grad_norm_
  scal[0] = norm;
  scal[1] = coef * grad_scale;
}
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        float* __restrict__ ema, size_t n, float lr, float beta1,
                                                        float beta2, float eps, float wd, float bc1, float rsqrt_bc2,
                                                        const float* __restrict__ scal, float ema_w, int ema_copy) {
  const float gs = scal[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * rsqrt_bc2 + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (ema != nullptr) ema[i] = ema_copy ? pi : fmaf(ema_w, pi - ema[i], ema[i]);  // lerp(ema, p, 1 - decay)
  }
}

int launch_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, size_t n, int step, float lr,
                     float beta1, float beta2, float eps, float wd, float max_norm, float grad_scale, float ema_decay,
                     int ema_copy, double* acc_scratch, float* scal_out, cudaStream_t s) {
  OSD_CHECK(p && g && m && v && acc_scratch && scal_out && n > 0 && step >= 1, "adamw_ema: bad arguments");
  OSD_CUDA(cudaMemsetAsync(acc_scratch, 0, sizeof(double), s));
  const int blocks = num_sms() * 4;
  sumsq_kernel<<<blocks, 256, 0, s>>>(g, n, acc_scratch);
  OSD_LAUNCHED();
  clip_coef_kernel<<<1, 1, 0, s>>>(acc_scratch, grad_scale, max_norm, scal_out);
  OSD_LAUNCHED();
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adamw_ema_kernel<<<blocks, 256, 0, s>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, wd, (float)bc1,
                                          (float)(1.0 / sqrt(bc2)), scal_out, 1.0f - ema_decay, ema_copy);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd

extern "C" {
__attribute__((visibility("default"))) int osd_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema,
                                                              size_t n, int step, float lr, float beta1, float beta2,
                                                              float eps, float weight_decay, float max_grad_norm,
                                                              float grad_scale, float ema_decay, int ema_copy,
                                                              double* acc_scratch, float* scal_out, void* stream) {
  return osd::launch_adamw_ema(p, g, m, v, ema, n, step, lr, beta1, beta2, eps, weight_decay, max_grad_norm, grad_scale,
                               ema_decay, ema_copy, acc_scratch, scal_out, static_cast<cudaStream_t>(stream));
}
}
