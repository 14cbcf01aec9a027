// Fused distance-marching loss of fit-denoiser (osu_dreamer/models/diffusion/train.py:86-108 of the reference),
// forward value + analytic gradients w.r.t. the model outputs (u_pred [B], v_pred [B,6,L]) in two passes over
// the [B,6,L] tensors:
//   d_sq[b]   = mean_l sum_e (xt - x1)^2            u_t = sqrt(d_sq + c0)
//   osl       = mean_b [ mean_l sum_e (xt - u v - x1)^2 / (d_sq + c0) ]
//   del       = mean_b [ mean_l sum_e (v - (xt - x1)/u_t)^2 ]
//   loss      = osl_w * osl + del_w * del ;   u_mape = mean_b |u - u_t| / u_t   (monitor only)
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

// acc layout (fp32, zeroed by the launcher): [0,B) d_sq sums | [B,2B) osl sums | [2B,3B) del sums | [3B,4B) du sums
__global__ void __launch_bounds__(256) loss_dsq_kernel(const float* __restrict__ xt, const float* __restrict__ x1,
                                                       float* __restrict__ acc, int per_sample) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const size_t base = (size_t)b * per_sample;
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const float r = xt[base + i] - x1[base + i];
    s = fmaf(r, r, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(acc + b, t);
  }
}
__global__ void __launch_bounds__(256) loss_main_kernel(const float* __restrict__ xt, const float* __restrict__ x1,
                                                        const float* __restrict__ u, const float* __restrict__ v,
                                                        float* __restrict__ acc, float* __restrict__ dv, float c0,
                                                        float osl_w, float del_w, int B, int L) {
  __shared__ float red[3][8];
  const int b = blockIdx.y;
  const int per_sample = 6 * L;
  const size_t base = (size_t)b * per_sample;
  const float dsq = acc[b] / (float)L;
  const float den = dsq + c0;
  const float ut = sqrtf(den);
  const float ub = u[b];
  const float k_osl = osl_w * 2.0f / ((float)B * (float)L * den);
  const float k_del = del_w * 2.0f / ((float)B * (float)L);
  const float inv_ut = 1.0f / ut;
  float s_osl = 0.f, s_del = 0.f, s_du = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const float r = xt[base + i] - x1[base + i];
    const float vv = v[base + i];
    const float e1 = r - ub * vv;       // denoised - x1
    const float e2 = vv - r * inv_ut;   // v_pred - v_target
    s_osl = fmaf(e1, e1, s_osl);
    s_del = fmaf(e2, e2, s_del);
    s_du = fmaf(e1, -vv, s_du);
    dv[base + i] = k_osl * (-ub) * e1 + k_del * e2;
  }
  s_osl = warp_sum(s_osl), s_del = warp_sum(s_del), s_du = warp_sum(s_du);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s_osl, red[1][threadIdx.x >> 5] = s_del, red[2][threadIdx.x >> 5] = s_du;
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(acc + (1 + threadIdx.x) * B + b, t);
  }
}
// out[0..3] = loss, osl, del, u_mape ; du[b] = d loss / d u_pred[b]
__global__ void loss_final_kernel(const float* __restrict__ acc, const float* __restrict__ u, float* __restrict__ out,
                                  float* __restrict__ du, float c0, float osl_w, float del_w, int B, int L) {
  if (threadIdx.x != 0) return;
  float osl = 0.f, del = 0.f, mape = 0.f;
  for (int b = 0; b < B; ++b) {
    const float den = acc[b] / (float)L + c0;
    const float ut = sqrtf(den);
    osl += acc[B + b] / ((float)L * den);
    del += acc[2 * B + b] / (float)L;
    mape += fabsf(u[b] - ut) / ut;
    du[b] = osl_w * 2.0f * acc[3 * B + b] / ((float)B * (float)L * den);
  }
  osl /= (float)B, del /= (float)B, mape /= (float)B;
  out[0] = osl_w * osl + del_w * del;
  out[1] = osl;
  out[2] = del;
  out[3] = mape;
}

int launch_loss(const float* xt, const float* x1, const float* u, const float* v, float c0, float osl_w, float del_w,
                int B, int L, float* out4, float* du, float* dv, float* acc, cudaStream_t s) {
  OSD_CHECK(xt && x1 && u && v && out4 && du && dv && acc && B > 0 && L > 0, "loss: bad arguments");
  OSD_CUDA(cudaMemsetAsync(acc, 0, (size_t)4 * B * sizeof(float), s));
  const int per_sample = 6 * L;
  int gx = ceil_div(per_sample, 256 * 4);
  if (gx > 64) gx = 64;
  dim3 grid(gx, B);
  loss_dsq_kernel<<<grid, 256, 0, s>>>(xt, x1, acc, per_sample);
  OSD_LAUNCHED();
  loss_main_kernel<<<grid, 256, 0, s>>>(xt, x1, u, v, acc, dv, c0, osl_w, del_w, B, L);
  OSD_LAUNCHED();
  loss_final_kernel<<<1, 32, 0, s>>>(acc, u, out4, du, c0, osl_w, del_w, B, L);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd

extern "C" {
__attribute__((visibility("default"))) int osd_loss_fwd_bwd(const float* xt, const float* x1, const float* u_pred,
                                                            const float* v_pred, float c0, float osl_weight,
                                                            float del_weight, int B, int L, float* out4, float* du,
                                                            float* dv, float* scratch, void* stream) {
  return osd::launch_loss(xt, x1, u_pred, v_pred, c0, osl_weight, del_weight, B, L, out4, du, dv, scratch,
                          static_cast<cudaStream_t>(stream));
}
}
