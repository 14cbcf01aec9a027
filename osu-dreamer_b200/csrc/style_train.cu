// fit-style on the B200 path: StyleTrainer.forward (osu_dreamer/models/style/train.py:48-91) and its backward in one call.
// The model (models/style/model.py:27-99) is a 6 M-parameter FiLM-modulated MLP on ONE 256-wide vector per sample, trained at
// batch 512 (models/style/model.yml:49): 18 GFLOP per step, i.e. launch latency, not throughput -- the reference issues
// ~1000 eager kernels per step.  Here the whole loss + gradient is ~200 launches from one C call, in exact fp32 on the CUDA
// cores (the tensor pipe would buy nothing at this size and fp32 keeps the 1e-3 parity class):
//   * one strided fp32 GEMM kernel serves every Linear: forward (X W^T + b), dgrad (dY W), wgrad (dY^T X, accumulating);
//   * row-wise kernels (warp per sample, 8 features per lane) for RMSNorm + FiLM, SiLU, gated residual, the heads and
//     their backward twins; column sums for the bias gradients; a fused loss forward + output gradient.
// Activations are saved in a caller-owned workspace (osd_style_train_workspace_floats).  Gradients ACCUMULATE into
// grads[60] (reference state-dict order; entries 3, 4 are the Fourier-feature buffers and are not touched).
#include "kernels.cuh"
#include "ptx.cuh"

#include <math.h>

namespace osd {

static constexpr int TS = 32, TH = 256, TE = 1024, TF = 128, TL = 5, TD = 8;
static constexpr float T_EPS = 1e-6f, T_EPS32 = 1.1920929e-07f;

// ------------------------------------------------------------------------------------------------ strided fp32 GEMM
// C[m, n] (+)= act(sum_k A(m, k) B(k, n) + bias[n]);  A(m, k) = A[m * sam + k * sak],  B(k, n) = B[k * sbk + n * sbn]
template <bool ACC>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long sam, long sak, const float* __restrict__ Bm,
                                                    long sbk, long sbn, float* __restrict__ C, long ldc,
                                                    const float* __restrict__ bias, int M, int N, int K) {
  __shared__ float As[16][68], Bs[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (sak == 1) { m = idx >> 4; k = idx & 15; } else { k = idx >> 6; m = idx & 63; }
      As[k][m] = (m0 + m < M && k0 + k < K) ? A[(long)(m0 + m) * sam + (long)(k0 + k) * sak] : 0.f;
      int n, kb;
      if (sbn == 1) { kb = idx >> 6; n = idx & 63; } else { n = idx >> 4; kb = idx & 15; }
      Bs[kb][n] = (n0 + n < N && k0 + kb < K) ? Bm[(long)(k0 + kb) * sbk + (long)(n0 + n) * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i], b[i] = Bs[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias != nullptr ? bias[n] : 0.f);
      float* c = C + (long)m * ldc + n;
      *c = ACC ? *c + v : v;
    }
  }
}
static int sgemm(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc, const float* bias,
                 int M, int N, int K, bool acc, cudaStream_t s) {
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  if (acc)
    sgemm_kernel<true><<<grid, 256, 0, s>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K);
  else
    sgemm_kernel<false><<<grid, 256, 0, s>>>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K);
  OSD_LAUNCHED();
  return 0;
}
// Y[B, N] = X[B, K] W[N, K]^T + b
static int lin_fwd(const float* X, const float* W, const float* b, float* Y, int Bn, int N, int K, cudaStream_t s) {
  return sgemm(X, K, 1, W, 1, K, Y, N, b, Bn, N, K, false, s);
}
// dX[B, K] (+)= dY[B, N] W[N, K]
static int lin_dgrad(const float* dY, const float* W, float* dX, int Bn, int N, int K, bool acc, cudaStream_t s) {
  return sgemm(dY, N, 1, W, K, 1, dX, K, nullptr, Bn, K, N, acc, s);
}
// dW[N, K] += dY[B, N]^T X[B, K]   (X rows may be strided: ldx)
static int lin_wgrad(const float* dY, long ldy, const float* X, long ldx, float* dW, int Bn, int N, int K, cudaStream_t s) {
  return sgemm(dY, 1, ldy, X, ldx, 1, dW, K, nullptr, N, K, Bn, true, s);
}

// out[n] += sum_b X[b * ld + n] * (mask == nullptr ? 1 : (mask[b * mstride] < 0) == want_neg)
__global__ void colsum_kernel(const float* __restrict__ X, long ld, int Bn, int N, float* __restrict__ out,
                              const float* __restrict__ mask, int mstride, int want_neg) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
  float a = 0.f;
  if (n < N)
    for (int b = r; b < Bn; b += 8) {
      if (mask != nullptr && ((mask[(long)b * mstride] < 0.f) != (want_neg != 0))) continue;
      a += X[(long)b * ld + n];
    }
  red[r][threadIdx.x & 31] = a;
  __syncthreads();
  if (r == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[n] += t;
  }
}
static int colsum(const float* X, long ld, int Bn, int N, float* out, cudaStream_t s, const float* mask = nullptr,
                  int mstride = 0, int want_neg = 0) {
  colsum_kernel<<<ceil_div(N, 32), 256, 0, s>>>(X, ld, Bn, N, out, mask, mstride, want_neg);
  OSD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------ row-wise kernels
// warp per sample; lane owns features [4 lane, +4) and [128 + 4 lane, +4)
struct Row8 {
  float v[8];
};
__device__ __forceinline__ Row8 ld8(const float* p, int lane) {
  Row8 r;
  const float4 a = *reinterpret_cast<const float4*>(p + lane * 4), b = *reinterpret_cast<const float4*>(p + 128 + lane * 4);
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st8(float* p, int lane, const Row8& r) {
  *reinterpret_cast<float4*>(p + lane * 4) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 128 + lane * 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ float dot8(const Row8& a, const Row8& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a.v[i], b.v[i], s);
  return warp_sum(s);
}
#define ROW_PROLOGUE                                              \
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);              \
  const int lane = threadIdx.x & 31;                              \
  if (b >= Bn) return

// Fourier features of the labels (fourier_features.py:15-16 on labels / 10, model.py:77), zeroed where the label is masked
// (< 0) so that the same buffer is the wgrad operand; feat [B, 5, 128]
__global__ void style_feat_kernel(const float* __restrict__ labels, const float* __restrict__ rW, const float* __restrict__ rb,
                                  float* __restrict__ feat, int Bn) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Bn * TL * TF) return;
  const int f = i % TF, bn = i / TF;
  const float lab = labels[bn];
  feat[i] = lab < 0.f ? 0.f : sqrtf(2.0f / TF) * cosf(lab / 10.0f * rW[f] + rb[f]);
}
// c[b] += sum_n (label < 0 ? null[n] : cond_b[n])   (the feat @ cond_w part is accumulated by 5 GEMMs)
__global__ void style_cond_bias_kernel(const float* __restrict__ labels, const float* __restrict__ cond_b,
                                       const float* __restrict__ null_l, float* __restrict__ c, int Bn) {
  ROW_PROLOGUE;
  Row8 acc = {};
  for (int n = 0; n < TL; ++n) {
    const Row8 t = ld8((labels[b * TL + n] < 0.f ? null_l : cond_b) + n * TH, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += t.v[i];
  }
  st8(c + (size_t)b * TH, lane, acc);
}
// h = rms_norm(x) (1 + scale) + shift   (model.py:91-92); mod [B, 768] = scale | shift | gate
__global__ void style_prenorm_kernel(const float* __restrict__ x, const float* __restrict__ mod, float* __restrict__ h, int Bn) {
  ROW_PROLOGUE;
  const Row8 xv = ld8(x + (size_t)b * TH, lane), sc = ld8(mod + (size_t)b * 3 * TH, lane), sh = ld8(mod + (size_t)b * 3 * TH + TH, lane);
  const float r = rsqrtf(dot8(xv, xv) * (1.0f / TH) + T_EPS);
  Row8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = xv.v[i] * r * (1.0f + sc.v[i]) + sh.v[i];
  st8(h + (size_t)b * TH, lane, o);
}
// given dh: dmod.scale = dh n, dmod.shift = dh, dx += r (dn - n mean(dn n)), dn = dh (1 + scale)
__global__ void style_prenorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mod, const float* __restrict__ dh,
                                         float* __restrict__ dx, float* __restrict__ dmod, int Bn) {
  ROW_PROLOGUE;
  const Row8 xv = ld8(x + (size_t)b * TH, lane), sc = ld8(mod + (size_t)b * 3 * TH, lane), g = ld8(dh + (size_t)b * TH, lane);
  const float r = rsqrtf(dot8(xv, xv) * (1.0f / TH) + T_EPS);
  Row8 n, dn, ds;
#pragma unroll
  for (int i = 0; i < 8; ++i) n.v[i] = xv.v[i] * r, dn.v[i] = g.v[i] * (1.0f + sc.v[i]), ds.v[i] = g.v[i] * n.v[i];
  const float mean = dot8(dn, n) * (1.0f / TH);
  Row8 o = ld8(dx + (size_t)b * TH, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] += r * (dn.v[i] - n.v[i] * mean);
  st8(dx + (size_t)b * TH, lane, o);
  st8(dmod + (size_t)b * 3 * TH, lane, ds);
  st8(dmod + (size_t)b * 3 * TH + TH, lane, g);
}
// x_out = x + rms_norm(h2) gate   (model.py:94-95)
__global__ void style_postnorm_kernel(const float* __restrict__ x, const float* __restrict__ h2, const float* __restrict__ mod,
                                      float* __restrict__ xo, int Bn) {
  ROW_PROLOGUE;
  const Row8 xv = ld8(x + (size_t)b * TH, lane), hv = ld8(h2 + (size_t)b * TH, lane), gt = ld8(mod + (size_t)b * 3 * TH + 2 * TH, lane);
  const float r = rsqrtf(dot8(hv, hv) * (1.0f / TH) + T_EPS);
  Row8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = xv.v[i] + hv.v[i] * r * gt.v[i];
  st8(xo + (size_t)b * TH, lane, o);
}
// given g = dx_out (dx keeps g: the skip path): dmod.gate = g n2, dh2 = r2 (dn2 - n2 mean(dn2 n2)), dn2 = g gate
__global__ void style_postnorm_bwd_kernel(const float* __restrict__ h2, const float* __restrict__ mod, const float* __restrict__ g_,
                                          float* __restrict__ dh2, float* __restrict__ dmod, int Bn) {
  ROW_PROLOGUE;
  const Row8 hv = ld8(h2 + (size_t)b * TH, lane), gt = ld8(mod + (size_t)b * 3 * TH + 2 * TH, lane), g = ld8(g_ + (size_t)b * TH, lane);
  const float r = rsqrtf(dot8(hv, hv) * (1.0f / TH) + T_EPS);
  Row8 n, dn, dg;
#pragma unroll
  for (int i = 0; i < 8; ++i) n.v[i] = hv.v[i] * r, dn.v[i] = g.v[i] * gt.v[i], dg.v[i] = g.v[i] * n.v[i];
  const float mean = dot8(dn, n) * (1.0f / TH);
  Row8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = r * (dn.v[i] - n.v[i] * mean);
  st8(dh2 + (size_t)b * TH, lane, o);
  st8(dmod + (size_t)b * 3 * TH + 2 * TH, lane, dg);
}
// act = silu(pre)  /  dpre = dact silu'(pre)   (in place on the second argument)
__global__ void style_silu_kernel(const float* __restrict__ pre, float* __restrict__ act, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) act[i] = pre[i] / (1.0f + expf(-pre[i]));
}
__global__ void style_silu_bwd_kernel(const float* __restrict__ pre, float* __restrict__ d, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    const float p = pre[i], sg = 1.0f / (1.0f + expf(-p));
    d[i] *= sg * (1.0f + p * (1.0f - sg));
  }
}
// heads (model.py:97-98): hn = RMSNorm_g(x) (eps = fp32 eps), z = u_out(rms_norm(x)), u = u_scale softplus(z)
__global__ void style_heads_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ wu,
                                   const float* __restrict__ bu, float* __restrict__ hn, float* __restrict__ z,
                                   float* __restrict__ u, int Bn) {
  ROW_PROLOGUE;
  const Row8 xv = ld8(x + (size_t)b * TH, lane), gv = ld8(g, lane), wv = ld8(wu, lane);
  const float ms = dot8(xv, xv) * (1.0f / TH);
  const float ro = rsqrtf(ms + T_EPS32), ru = rsqrtf(ms + T_EPS);
  Row8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = xv.v[i] * ro * gv.v[i];
  st8(hn + (size_t)b * TH, lane, o);
  const float zz = dot8(xv, wv) * ru + bu[0];
  if (lane == 0) {
    z[b] = zz;
    u[b] = sqrtf(2.0f * TS) * (zz > 20.f ? zz : log1pf(expf(zz)));
  }
}
// given dhn [B,256] (= dv Wo) and du [B]: dx = both head paths; dg += colsum(dhn x ro) and dwu / dbu through scratch rows:
// gpart [B,256] = dhn x ro, upart [B,256] = dz x ru, dzv [B] = dz
__global__ void style_heads_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ wu,
                                       const float* __restrict__ z, const float* __restrict__ dhn, const float* __restrict__ du,
                                       float* __restrict__ dx, float* __restrict__ gpart, float* __restrict__ upart,
                                       float* __restrict__ dzv, int Bn) {
  ROW_PROLOGUE;
  const Row8 xv = ld8(x + (size_t)b * TH, lane), gv = ld8(g, lane), wv = ld8(wu, lane), dh = ld8(dhn + (size_t)b * TH, lane);
  const float ms = dot8(xv, xv) * (1.0f / TH);
  const float ro = rsqrtf(ms + T_EPS32), ru = rsqrtf(ms + T_EPS);
  const float zz = z[b];
  const float dz = du[b] * sqrtf(2.0f * TS) * (zz > 20.f ? 1.0f : 1.0f / (1.0f + expf(-zz)));
  Row8 no, nu, dno, dnu, gp, up;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    no.v[i] = xv.v[i] * ro, nu.v[i] = xv.v[i] * ru;
    dno.v[i] = dh.v[i] * gv.v[i], dnu.v[i] = dz * wv.v[i];
    gp.v[i] = dh.v[i] * no.v[i], up.v[i] = dz * nu.v[i];
  }
  const float mo = dot8(dno, no) * (1.0f / TH), mu = dot8(dnu, nu) * (1.0f / TH);
  Row8 o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = ro * (dno.v[i] - no.v[i] * mo) + ru * (dnu.v[i] - nu.v[i] * mu);
  st8(dx + (size_t)b * TH, lane, o);
  st8(gpart + (size_t)b * TH, lane, gp);
  st8(upart + (size_t)b * TH, lane, up);
  if (lane == 0) dzv[b] = dz;
}
// StyleTrainer.forward's loss (train.py:70-88) and its gradients w.r.t. u_pred, v_pred; acc4 (zeroed) -> sums
__global__ void style_loss_kernel(const float* __restrict__ st, const float* __restrict__ s1, const float* __restrict__ u,
                                  const float* __restrict__ v, float c0, float osl_w, float del_w, int Bn, float* __restrict__ acc4,
                                  float* __restrict__ du, float* __restrict__ dv) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  float osl = 0.f, del = 0.f, mape = 0.f;
  if (b < Bn) {
    const float x = st[b * TS + lane], y = s1[b * TS + lane], vv = v[b * TS + lane], uu = u[b];
    const float diff = x - y;
    const float dsq = warp_sum(diff * diff);
    const float den = dsq + c0, ut = sqrtf(den);
    const float e1 = x - uu * vv - y;          // denoised - s1
    const float e2 = vv - diff / ut;           // v_pred - v_target
    osl = warp_sum(e1 * e1) / den;
    del = warp_sum(e2 * e2);
    mape = fabsf(uu - ut) / ut;
    const float inv_b = 1.0f / (float)Bn;
    dv[b * TS + lane] = inv_b * (osl_w * 2.0f * e1 * (-uu) / den + del_w * 2.0f * e2);
    const float dus = warp_sum(e1 * vv);
    if (lane == 0) du[b] = inv_b * osl_w * (-2.0f) * dus / den;
  }
  if (lane == 0 && b < Bn) {
    atomicAdd(acc4 + 1, osl);
    atomicAdd(acc4 + 2, del);
    atomicAdd(acc4 + 3, mape);
  }
}
__global__ void style_loss_final_kernel(const float* __restrict__ acc4, float osl_w, float del_w, int Bn, float* __restrict__ out4) {
  const float osl = acc4[1] / Bn, del = acc4[2] / Bn;
  out4[0] = osl_w * osl + del_w * del;
  out4[1] = osl;
  out4[2] = del;
  out4[3] = acc4[3] / Bn;
}

// ------------------------------------------------------------------------------------------------ workspace
struct StylePlan {
  size_t feat, c, mod, x, hb, pre, h2, act, hn, z, dx, dh2, dact, dhb, dmod, dc, gpart, upart, dzv, total;
};
static StylePlan style_plan(int B) {
  StylePlan p;
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t r = o;
    o += (n + 63) / 64 * 64;
    return r;
  };
  const size_t b = B;
  p.feat = take(b * TL * TF);
  p.c = take(b * TH);
  p.mod = take(TD * b * 3 * TH);
  p.x = take((TD + 1) * b * TH);
  p.hb = take(TD * b * TH);
  p.pre = take(TD * b * TE);
  p.h2 = take(TD * b * TH);
  p.act = take(b * TE);
  p.hn = take(b * TH);
  p.z = take(b);
  p.dx = take(b * TH);
  p.dh2 = take(b * TH);
  p.dact = take(b * TE);
  p.dhb = take(b * TH);
  p.dmod = take(b * 3 * TH);
  p.dc = take(b * TH);
  p.gpart = take(b * TH);
  p.upart = take(b * TH);
  p.dzv = take(b);
  p.total = o;
  return p;
}
size_t style_train_workspace_floats(int B) { return style_plan(B).total; }

// parameter indices (reference state-dict order, include/osd_b200.h OSD_STYLE_NUM_PARAMS)
enum { SP_CW = 0, SP_CB, SP_NULL, SP_RW, SP_RB, SP_INW, SP_INB, SP_OG, SP_OW, SP_OB, SP_UW, SP_UB, SP_FILM0 = 12, SP_BLK0 = 28 };

// ---- forward with saved activations (model.py:72-99): st [B,32], labels [B,5] (< 0 = masked) -> u [B], v [B,32]
int launch_style_train_forward(const float* const* P, const float* st, const float* labels, float* u_out, float* v_out,
                               float* ws, int B, cudaStream_t s) {
  OSD_CHECK(P && st && labels && u_out && v_out && ws && B > 0, "style_train_forward: bad arguments");
  const StylePlan pl = style_plan(B);
  const int rows = ceil_div(B, 8);
  float *feat = ws + pl.feat, *c = ws + pl.c, *mod = ws + pl.mod, *x = ws + pl.x, *hb = ws + pl.hb, *pre = ws + pl.pre,
        *h2 = ws + pl.h2, *act = ws + pl.act, *hn = ws + pl.hn, *z = ws + pl.z;
  auto X = [&](int i) { return x + (size_t)i * B * TH; };
  auto MOD = [&](int i) { return mod + (size_t)i * B * 3 * TH; };
  const size_t nE = (size_t)B * TE;
  style_feat_kernel<<<ceil_div(B * TL * TF, 256), 256, 0, s>>>(labels, P[SP_RW], P[SP_RB], feat, B);
  OSD_LAUNCHED();
  style_cond_bias_kernel<<<rows, 256, 0, s>>>(labels, P[SP_CB], P[SP_NULL], c, B);
  OSD_LAUNCHED();
  for (int n = 0; n < TL; ++n)  // c += feat[:, n, :] cond_w[n]  ([128, 256] row-major = B(k, n) with sbk = 256)
    OSD_TRY(sgemm(feat + n * TF, TL * TF, 1, P[SP_CW] + (size_t)n * TF * TH, TH, 1, c, TH, nullptr, B, TH, TF, true, s));
  OSD_TRY(lin_fwd(st, P[SP_INW], P[SP_INB], X(0), B, TH, TS, s));
  for (int i = 0; i < TD; ++i) {
    const float *w0 = P[SP_BLK0 + 4 * i], *b0 = P[SP_BLK0 + 4 * i + 1], *w3 = P[SP_BLK0 + 4 * i + 2], *b3 = P[SP_BLK0 + 4 * i + 3];
    OSD_TRY(lin_fwd(c, P[SP_FILM0 + 2 * i], P[SP_FILM0 + 2 * i + 1], MOD(i), B, 3 * TH, TH, s));
    style_prenorm_kernel<<<rows, 256, 0, s>>>(X(i), MOD(i), hb + (size_t)i * B * TH, B);
    OSD_LAUNCHED();
    OSD_TRY(lin_fwd(hb + (size_t)i * B * TH, w0, b0, pre + (size_t)i * nE, B, TE, TH, s));
    style_silu_kernel<<<(unsigned)((nE + 255) / 256), 256, 0, s>>>(pre + (size_t)i * nE, act, nE);
    OSD_LAUNCHED();
    OSD_TRY(lin_fwd(act, w3, b3, h2 + (size_t)i * B * TH, B, TH, TE, s));
    style_postnorm_kernel<<<rows, 256, 0, s>>>(X(i), h2 + (size_t)i * B * TH, MOD(i), X(i + 1), B);
    OSD_LAUNCHED();
  }
  style_heads_kernel<<<rows, 256, 0, s>>>(X(TD), P[SP_OG], P[SP_UW], P[SP_UB], hn, z, u_out, B);
  OSD_LAUNCHED();
  OSD_TRY(lin_fwd(hn, P[SP_OW], P[SP_OB], v_out, B, TS, TH, s));
  return 0;
}

// ---- loss of StyleTrainer.forward (train.py:70-88) on [B,32] vectors: out4 = {loss, osl, del, u_mape}, du [B], dv [B,32];
//      acc_scratch: 4 floats
int launch_style_loss(const float* st, const float* s1, const float* u, const float* v, float osl_w, float del_w, float* out4,
                      float* du, float* dv, float* acc_scratch, int B, cudaStream_t s) {
  OSD_CHECK(st && s1 && u && v && out4 && du && dv && acc_scratch && B > 0, "style_loss: bad arguments");
  const float d0_sq = 2.0f * TS;
  const float t99 = 1.0f / (1.0f + expf(-2.3263478740408408f));
  const float c0 = (1.0f - t99) * (1.0f - t99) * d0_sq;  // model.py:34-39
  OSD_CUDA(cudaMemsetAsync(acc_scratch, 0, 4 * sizeof(float), s));
  style_loss_kernel<<<ceil_div(B, 8), 256, 0, s>>>(st, s1, u, v, c0, osl_w, del_w, B, acc_scratch, du, dv);
  OSD_LAUNCHED();
  style_loss_final_kernel<<<1, 1, 0, s>>>(acc_scratch, osl_w, del_w, B, out4);
  OSD_LAUNCHED();
  return 0;
}

// ---- backward of launch_style_train_forward: given du [B], dv [B,32] ACCUMULATES the parameter gradients into G[60]
int launch_style_backward(const float* const* P, const float* st, const float* labels, const float* du, const float* dv,
                          float* const* G, float* ws, int B, cudaStream_t s) {
  OSD_CHECK(P && st && labels && du && dv && G && ws && B > 0, "style_backward: bad arguments");
  const StylePlan pl = style_plan(B);
  const int rows = ceil_div(B, 8);
  float *feat = ws + pl.feat, *c = ws + pl.c, *mod = ws + pl.mod, *x = ws + pl.x, *hb = ws + pl.hb, *pre = ws + pl.pre,
        *h2 = ws + pl.h2, *act = ws + pl.act, *hn = ws + pl.hn, *z = ws + pl.z;
  auto X = [&](int i) { return x + (size_t)i * B * TH; };
  auto MOD = [&](int i) { return mod + (size_t)i * B * 3 * TH; };
  const size_t nE = (size_t)B * TE;
  float *dx = ws + pl.dx, *dh2 = ws + pl.dh2, *dact = ws + pl.dact, *dhb = ws + pl.dhb, *dmod = ws + pl.dmod, *dc = ws + pl.dc,
        *gpart = ws + pl.gpart, *upart = ws + pl.upart, *dzv = ws + pl.dzv;
  float* dhn = dhb;  // [B,256] scratch, free until the first block
  OSD_TRY(lin_dgrad(dv, P[SP_OW], dhn, B, TS, TH, false, s));
  OSD_TRY(lin_wgrad(dv, TS, hn, TH, G[SP_OW], B, TS, TH, s));
  OSD_TRY(colsum(dv, TS, B, TS, G[SP_OB], s));
  style_heads_bwd_kernel<<<rows, 256, 0, s>>>(X(TD), P[SP_OG], P[SP_UW], z, dhn, du, dx, gpart, upart, dzv, B);
  OSD_LAUNCHED();
  OSD_TRY(colsum(gpart, TH, B, TH, G[SP_OG], s));
  OSD_TRY(colsum(upart, TH, B, TH, G[SP_UW], s));
  OSD_TRY(colsum(dzv, 1, B, 1, G[SP_UB], s));
  OSD_CUDA(cudaMemsetAsync(dc, 0, (size_t)B * TH * 4, s));
  for (int i = TD - 1; i >= 0; --i) {
    const float *w0 = P[SP_BLK0 + 4 * i], *w3 = P[SP_BLK0 + 4 * i + 2];
    float *gw0 = G[SP_BLK0 + 4 * i], *gb0 = G[SP_BLK0 + 4 * i + 1], *gw3 = G[SP_BLK0 + 4 * i + 2], *gb3 = G[SP_BLK0 + 4 * i + 3];
    style_postnorm_bwd_kernel<<<rows, 256, 0, s>>>(h2 + (size_t)i * B * TH, MOD(i), dx, dh2, dmod, B);
    OSD_LAUNCHED();
    style_silu_kernel<<<(unsigned)((nE + 255) / 256), 256, 0, s>>>(pre + (size_t)i * nE, act, nE);  // recomputed, not saved
    OSD_LAUNCHED();
    OSD_TRY(lin_wgrad(dh2, TH, act, TE, gw3, B, TH, TE, s));
    OSD_TRY(colsum(dh2, TH, B, TH, gb3, s));
    OSD_TRY(lin_dgrad(dh2, w3, dact, B, TH, TE, false, s));
    style_silu_bwd_kernel<<<(unsigned)((nE + 255) / 256), 256, 0, s>>>(pre + (size_t)i * nE, dact, nE);
    OSD_LAUNCHED();
    OSD_TRY(lin_wgrad(dact, TE, hb + (size_t)i * B * TH, TH, gw0, B, TE, TH, s));
    OSD_TRY(colsum(dact, TE, B, TE, gb0, s));
    OSD_TRY(lin_dgrad(dact, w0, dhb, B, TE, TH, false, s));
    style_prenorm_bwd_kernel<<<rows, 256, 0, s>>>(X(i), MOD(i), dhb, dx, dmod, B);
    OSD_LAUNCHED();
    // film_i: mod = c Wf^T + bf
    OSD_TRY(lin_wgrad(dmod, 3 * TH, c, TH, G[SP_FILM0 + 2 * i], B, 3 * TH, TH, s));
    OSD_TRY(colsum(dmod, 3 * TH, B, 3 * TH, G[SP_FILM0 + 2 * i + 1], s));
    OSD_TRY(lin_dgrad(dmod, P[SP_FILM0 + 2 * i], dc, B, 3 * TH, TH, true, s));
  }
  // proj_in
  OSD_TRY(lin_wgrad(dx, TH, st, TS, G[SP_INW], B, TH, TS, s));
  OSD_TRY(colsum(dx, TH, B, TH, G[SP_INB], s));
  // conditioning (model.py:72-79): masked labels feed the null embeddings, the others cond_w / cond_b
  for (int n = 0; n < TL; ++n) {
    OSD_TRY(sgemm(feat + n * TF, 1, TL * TF, dc, TH, 1, G[SP_CW] + (size_t)n * TF * TH, TH, nullptr, TF, TH, B, true, s));
    OSD_TRY(colsum(dc, TH, B, TH, G[SP_CB] + n * TH, s, labels + n, TL, 0));
    OSD_TRY(colsum(dc, TH, B, TH, G[SP_NULL] + n * TH, s, labels + n, TL, 1));
  }
  return 0;
}

}  // namespace osd
