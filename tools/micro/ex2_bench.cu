// micro-benchmark: SFU throughput of ex2.approx.ftz.f32 vs ex2.approx.ftz.bf16x2 vs polynomial emulation (FFMA2)
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__global__ void k_f32(float* out, float x0, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = x0 + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16x2(float* out, float x0, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) { __nv_bfloat162 v = __floats2bfloat162_rn(x0 + i, x0 - i); a[i] = *(unsigned*)&v; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
__global__ void k_f16x2(float* out, float x0, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003800u + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 3; ++which) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_f32<<<148 * 4, 256>>>(out, -0.5f, iters);
      if (which == 1) k_bf16x2<<<148 * 4, 256>>>(out, -0.5f, iters);
      if (which == 2) k_f16x2<<<148 * 4, 256>>>(out, -0.5f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = 148.0 * 4 * 256 * 8.0 * iters;  // thread-level ex2 instructions
      if (rep) printf("%s: %.3f ms, %.1f G thread-instr/s, per SM per clk (1.9GHz): %.2f\n",
                      which == 0 ? "ex2.f32" : which == 1 ? "ex2.bf16x2" : "ex2.f16x2", ms, instr / ms / 1e6,
                      instr / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
