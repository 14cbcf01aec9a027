// Backward of the distance head (osu_dreamer/models/diffusion/model.py:58-68, 99-102 of the reference):
//   f = mean_l silu(W4 dwconv3(silu(W1 dwconv3(xt) + b1)) + b4);  fm = f (1 + scale) + shift;
//   u = u_scale * softplus(wout . fm + bout)
// The per-token intermediates are recomputed from xt inside the block (halo 2), so nothing is saved.
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

struct UHeadW {
  const float *w0, *b0, *w1, *b1, *w3, *b3, *w4, *b4;
};
struct UHeadG {
  float *w0, *b0, *w1, *b1, *w3, *b3, *w4, *b4;
};

__device__ __forceinline__ float dsilu(float z) {
  const float sg = 1.0f / (1.0f + __expf(-z));
  return sg * (1.0f + z * (1.0f - sg));
}

// u = u_scale * softplus(zv) backward for the tiny [B,64] tail.  One block.
//   outputs: dfsum [B,64] (gradient w.r.t. the SUM over l, i.e. includes 1/L), dumod [B,128], dwout[64], dbout[1]
__global__ void u_final_bwd_kernel(const float* __restrict__ du, const float* __restrict__ fsum,
                                   const float* __restrict__ umod, const float* __restrict__ wout,
                                   const float* __restrict__ bout, float u_scale, float inv_L, float* __restrict__ dfsum,
                                   float* __restrict__ dumod, float* __restrict__ dwout, float* __restrict__ dbout,
                                   int Bn) {
  __shared__ float red[64];
  const int i = threadIdx.x;  // 64 threads
  float aw = 0.f, ab = 0.f;
  for (int b = 0; b < Bn; ++b) {
    const float f = fsum[(size_t)b * 64 + i] * inv_L;
    const float sc = umod[(size_t)b * 128 + i], sh = umod[(size_t)b * 128 + 64 + i];
    const float fm = f * (1.f + sc) + sh;
    red[i] = wout[i] * fm;
    __syncthreads();
    for (int o = 32; o > 0; o >>= 1) {
      if (i < o) red[i] += red[i + o];
      __syncthreads();
    }
    const float zv = red[0] + bout[0];
    __syncthreads();
    const float sg = zv > 20.f ? 1.f : 1.0f / (1.0f + expf(-zv));
    const float dz = du[b] * u_scale * sg;
    aw += dz * fm;
    ab += dz;
    const float dfm = dz * wout[i];
    dumod[(size_t)b * 128 + i] = dfm * f;
    dumod[(size_t)b * 128 + 64 + i] = dfm;
    dfsum[(size_t)b * 64 + i] = dfm * (1.f + sc) * inv_L;
  }
  dwout[i] += aw;
  if (i == 0) dbout[0] += ab;
}
int launch_u_final_bwd(const float* du, const float* fsum, const float* umod, const float* wout, const float* bout,
                       float u_scale, int L, float* dfsum, float* dumod, float* dwout, float* dbout, int B,
                       cudaStream_t s) {
  u_final_bwd_kernel<<<1, 64, 0, s>>>(du, fsum, umod, wout, bout, u_scale, 1.0f / (float)L, dfsum, dumod, dwout, dbout,
                                      B);
  OSD_LAUNCHED();
  return 0;
}

static constexpr int UB_TOK = 32;  // tokens per sub-tile
static constexpr int UB_SUB = 8;   // sub-tiles per block
// smem (floats): xs[38][6] c1[36][6] pre1[36][64] c2[34][64] pre2->dh2[34][64] dc2[34][64] dpre1[32][64] dc1[32][6] w4[64][65]
__global__ void __launch_bounds__(256) u_head_bwd_kernel(const float* __restrict__ xt, UHeadW w,
                                                         const float* __restrict__ dfsum, UHeadG g, int L) {
  extern __shared__ float sm[];
  float* xs = sm;                   // [38][6]   positions l0-3 .. l0+34
  float* c1 = xs + 38 * 6;          // [36][6]   positions l0-2 .. l0+33
  float* pre1 = c1 + 36 * 6;        // [36][64]
  float* c2 = pre1 + 36 * 64;       // [34][64]  positions l0-1 .. l0+32
  float* dh2 = c2 + 34 * 64;        // [34][64]  (pre2, then dh2 in place)
  float* dc2 = dh2 + 34 * 64;       // [34][64]
  float* dp1 = dc2 + 34 * 64;       // [32][64]  dpre1 at tile positions
  float* dc1 = dp1 + 32 * 64;       // [32][6]
  float* w4s = dc1 + 32 * 6;        // [64][65]
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * 64; i += 256) w4s[(i >> 6) * 65 + (i & 63)] = w.w4[i];
  // register accumulators
  float aW4[16];  // thread owns (i = tid >> 2 (0..63), k = (tid & 3) * 16 + j)
#pragma unroll
  for (int j = 0; j < 16; ++j) aW4[j] = 0.f;
  float ab4 = 0.f;                      // tid < 64: channel tid
  float aw3[3] = {0, 0, 0}, ab3 = 0.f;  // tid < 64: channel tid
  float aW1[6] = {0, 0, 0, 0, 0, 0}, ab1 = 0.f;  // tid < 64: row u = tid
  float aw0[3] = {0, 0, 0}, ab0 = 0.f;           // tid < 6: channel tid
  const float df_b[1] = {0};
  (void)df_b;

  for (int sub = 0; sub < UB_SUB; ++sub) {
    const int l0 = (blockIdx.x * UB_SUB + sub) * UB_TOK;
    if (l0 >= L) break;
    __syncthreads();
    for (int i = tid; i < 38 * 6; i += 256) {
      const int rr = i / 6, e = i % 6;
      const int l = l0 + rr - 3;
      xs[i] = (l >= 0 && l < L) ? xt[((size_t)b * 6 + e) * L + l] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < 36 * 6; i += 256) {
      const int rr = i / 6, e = i % 6;
      const int l = l0 + rr - 2;
      float acc = 0.f;
      if (l >= 0 && l < L)
        acc = w.b0[e] + w.w0[e * 3] * xs[rr * 6 + e] + w.w0[e * 3 + 1] * xs[(rr + 1) * 6 + e] +
              w.w0[e * 3 + 2] * xs[(rr + 2) * 6 + e];
      c1[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < 36 * 64; i += 256) {
      const int rr = i >> 6, u = i & 63;
      float acc = w.b1[u];
#pragma unroll
      for (int e = 0; e < 6; ++e) acc = fmaf(w.w1[u * 6 + e], c1[rr * 6 + e], acc);
      pre1[i] = acc;
    }
    __syncthreads();
    // c2 at positions l0-1 .. l0+32 ; h1 is zero outside the sequence (conv zero padding)
    for (int i = tid; i < 34 * 64; i += 256) {
      const int rr = i >> 6, u = i & 63;
      const int l = l0 + rr - 1;
      float acc = w.b3[u];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ll = l + j - 1;
        if (ll >= 0 && ll < L) acc = fmaf(w.w3[u * 3 + j], silu_f(pre1[(rr + j) * 64 + u]), acc);
      }
      c2[i] = acc;
    }
    __syncthreads();
    // dh2 = dfsum * silu'(pre2) for valid positions, else 0
    for (int i = tid; i < 34 * 64; i += 256) {
      const int rr = i >> 6, u = i & 63;
      const int l = l0 + rr - 1;
      float v = 0.f;
      if (l >= 0 && l < L) {
        float acc = w.b4[u];
#pragma unroll 16
        for (int k = 0; k < 64; ++k) acc = fmaf(w4s[u * 65 + k], c2[rr * 64 + k], acc);
        v = dfsum[(size_t)b * 64 + u] * dsilu(acc);
      }
      dh2[i] = v;
    }
    __syncthreads();
    // dc2[p][k] = sum_i W4[i][k] dh2[p][i]
    for (int i = tid; i < 34 * 64; i += 256) {
      const int rr = i >> 6, k = i & 63;
      float acc = 0.f;
#pragma unroll 16
      for (int u = 0; u < 64; ++u) acc = fmaf(w4s[u * 65 + k], dh2[rr * 64 + u], acc);
      dc2[i] = acc;
    }
    // dW4[i][k] += sum_{p in tile} dh2[p][i] c2[p][k] ; db4
    {
      const int i4 = tid >> 2, kb = (tid & 3) * 16;
      for (int r = 0; r < UB_TOK; ++r) {
        if (l0 + r >= L) break;
        const float d = dh2[(r + 1) * 64 + i4];
#pragma unroll
        for (int j = 0; j < 16; ++j) aW4[j] = fmaf(d, c2[(r + 1) * 64 + kb + j], aW4[j]);
      }
      if (tid < 64)
        for (int r = 0; r < UB_TOK; ++r) {
          if (l0 + r >= L) break;
          ab4 += dh2[(r + 1) * 64 + tid];
        }
    }
    __syncthreads();
    // dw3 / db3 over tile positions; dpre1 at tile positions
    if (tid < 64) {
      for (int r = 0; r < UB_TOK; ++r) {
        const int l = l0 + r;
        if (l >= L) break;
        const float d = dc2[(r + 1) * 64 + tid];
        ab3 += d;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int ll = l + j - 1;
          if (ll >= 0 && ll < L) aw3[j] = fmaf(d, silu_f(pre1[(r + 1 + j) * 64 + tid]), aw3[j]);
        }
      }
    }
    for (int i = tid; i < UB_TOK * 64; i += 256) {
      const int r = i >> 6, u = i & 63;
      const int l = l0 + r;
      float v = 0.f;
      if (l < L) {
        // dh1[l][u] = sum_j w3[u][j] dc2[l - j + 1][u]   (dc2 row index = position - (l0 - 1))
        float dh1 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int pp = l - j + 1;
          if (pp >= 0 && pp < L) dh1 = fmaf(w.w3[u * 3 + j], dc2[(pp - l0 + 1) * 64 + u], dh1);
        }
        v = dh1 * dsilu(pre1[(r + 2) * 64 + u]);
      }
      dp1[i] = v;
    }
    __syncthreads();
    if (tid < 64) {
      for (int r = 0; r < UB_TOK; ++r) {
        if (l0 + r >= L) break;
        const float d = dp1[r * 64 + tid];
        ab1 += d;
#pragma unroll
        for (int e = 0; e < 6; ++e) aW1[e] = fmaf(d, c1[(r + 2) * 6 + e], aW1[e]);
      }
    }
    for (int i = tid; i < UB_TOK * 6; i += 256) {
      const int r = i / 6, e = i % 6;
      float acc = 0.f;
      if (l0 + r < L)
        for (int u = 0; u < 64; ++u) acc = fmaf(w.w1[u * 6 + e], dp1[r * 64 + u], acc);
      dc1[i] = acc;
    }
    __syncthreads();
    if (tid < 6) {
      for (int r = 0; r < UB_TOK; ++r) {
        if (l0 + r >= L) break;
        const float d = dc1[r * 6 + tid];
        ab0 += d;
#pragma unroll
        for (int j = 0; j < 3; ++j) aw0[j] = fmaf(d, xs[(r + 2 + j) * 6 + tid], aw0[j]);
      }
    }
  }
  {
    const int i4 = tid >> 2, kb = (tid & 3) * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) atomicAdd(g.w4 + i4 * 64 + kb + j, aW4[j]);
  }
  if (tid < 64) {
    atomicAdd(g.b4 + tid, ab4);
    atomicAdd(g.b3 + tid, ab3);
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(g.w3 + tid * 3 + j, aw3[j]);
    atomicAdd(g.b1 + tid, ab1);
#pragma unroll
    for (int e = 0; e < 6; ++e) atomicAdd(g.w1 + tid * 6 + e, aW1[e]);
  }
  if (tid < 6) {
    atomicAdd(g.b0 + tid, ab0);
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(g.w0 + tid * 3 + j, aw0[j]);
  }
}
int launch_u_head_bwd(const float* xt, const float* const* w8, const float* dfsum, float* const* g8, int B, int L,
                      cudaStream_t s) {
  UHeadW w{w8[0], w8[1], w8[2], w8[3], w8[4], w8[5], w8[6], w8[7]};
  UHeadG g{g8[0], g8[1], g8[2], g8[3], g8[4], g8[5], g8[6], g8[7]};
  const int smem = (38 * 6 + 36 * 6 + 36 * 64 + 3 * 34 * 64 + 32 * 64 + 32 * 6 + 64 * 65) * 4;
  static DeviceOnce once;
  if (once.first()) {
    OSD_CUDA(cudaFuncSetAttribute(u_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid(ceil_div(L, UB_TOK * UB_SUB), B);
  u_head_bwd_kernel<<<grid, 256, smem, s>>>(xt, w, dfsum, g, L);
  OSD_LAUNCHED();
  return 0;
}

}  // namespace osd
