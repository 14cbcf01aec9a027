// micro-benchmark: how many softmax elements per clock one SM sustains as a function of the number of resident warps, for
// the exact per-chunk routine of the forward attention kernels (scale FFMA2 -> ex2 (SFU, or a share on the FMA pipe) ->
// row-sum FADD2 -> bf16 pack), registers only (no TMEM, no MMA).  Answers: is the softmax of attn_fwd_* bound by the SFU
// (16 ex2 / clk / SM), by instruction issue, or by thread-level parallelism (2 softmax warps per sub-partition)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../osu-dreamer_b200/csrc -o softmax_bench softmax_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace osd;

__device__ __forceinline__ float b_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MODE: 0 full routine emu 0, 1 emu 1 (quarter on FMA pipe), 2 emu 2 (half), 3 "direct": ex2 + pack only (pre-scaled S, sums by MMA)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_softmax(unsigned* out, float x0, int iters, long long* clk) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(x0 + 1e-3f * threadIdx.x + 0.01f * i);
  float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
  const float c = 0.18f, neg_mc = -3.f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
    const float2 c2 = make_float2(c, c), n2 = make_float2(neg_mc, neg_mc);
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      float2 e;
      if (MODE == 3) {
        e = make_float2(b_ex2(__uint_as_float(r[2 * p])), b_ex2(__uint_as_float(r[2 * p + 1])));
      } else {
        const float2 a = ffma2(make_float2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1])), c2, n2);
        if ((p & 3) < MODE) e = ex2_poly2(a); else e = make_float2(b_ex2(a.x), b_ex2(a.y));
        if (p & 1) s23 = fadd2(s23, e); else s01 = fadd2(s01, e);
      }
      pk[p] = pack_bf16(e.x, e.y);
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {  // feed the results back so nothing is hoisted or removed (stands in for the TMEM traffic)
      acc ^= pk[p];
      r[2 * p] = (r[2 * p] & 0xfffff000u) | (pk[p] & 0xfffu);
      r[2 * p + 1] = (r[2 * p + 1] & 0xfffff000u) | ((pk[p] >> 16) & 0xfffu);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(s01.x + s01.y + s23.x + s23.y);
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE>
void run(const char* name, unsigned* out, long long* clk) {
  for (int warps : {4, 8, 12, 16, 32}) {
    const int iters = 2048;
    k_softmax<MODE><<<148, warps * 32>>>(out, -0.5f, iters, clk);
    cudaDeviceSynchronize();
    k_softmax<MODE><<<148, warps * 32>>>(out, -0.5f, iters, clk);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double elems = (double)warps * 32 * 32 * iters;  // per SM
    printf("%-28s warps/SM %2d: %.2f elements/clk/SM (%lld clk)\n", name, warps, elems / (double)h, h);
  }
}
int main() {
  unsigned* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 8);
  run<0>("scale+ex2+sum+pack, emu 0", out, clk);
  run<1>("scale+ex2+sum+pack, emu 1/4", out, clk);
  run<2>("scale+ex2+sum+pack, emu 1/2", out, clk);
  run<3>("ex2+pack only (direct)", out, clk);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
