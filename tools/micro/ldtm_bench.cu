// micro-benchmark: TMEM -> register bandwidth of tcgen05.ld (32x32b): 32 cells / 32 registers per thread (x32),
// 64 cells / 64 registers (x64), and 64 cells packed into 32 registers (x32.pack::16b: two adjacent 16-bit columns per
// register).  Tells whether the 64 B/clk/SM limit counts TMEM cells read or register bytes written, i.e. whether
// fp16 accumulators read with .pack::16b would halve the attention kernels' TMEM-read floor.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void ld_x32_pack(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// mode 0: x32 (32 cells -> 32 regs); 1: two x32 at adjacent column blocks (64 cells -> 64 regs); 2: x32.pack::16b (64 cells -> 32 regs)
__global__ void __launch_bounds__(512, 1) k(int mode, int iters, long long* clk_out, uint32_t* sink, uint32_t* dump) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64;  // warps w, w+4, .. share a lane quadrant
  uint32_t r[32], q[32];
  for (int i = 0; i < 32; ++i) r[i] = 0x10000u * (2 * i + 1) + (uint32_t)(threadIdx.x);  // hi half = 2i+1, lo half = thread
  st_x32(base, r);
  for (int i = 0; i < 32; ++i) r[i] = 0x10000u * (2 * i + 65) + (uint32_t)(1000 + i);
  st_x32(base + 32, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  if (dump != nullptr && blockIdx.x == 0) {  // what does a packed load return?
    ld_x32_pack(base, q);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (threadIdx.x == 5) for (int i = 0; i < 32; ++i) dump[i] = q[i];
  }
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      ld_x32(base + (it & 1) * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[13] ^ r[31];
    } else if (mode == 1) {
      ld_x32(base, r);
      ld_x32(base + 32, q);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[13] ^ r[31] ^ q[0] ^ q[17] ^ q[31];
    } else {
      ld_x32_pack(base, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[13] ^ r[31];
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * 512 + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
  long long* clk; uint32_t *sink, *dump;
  cudaMalloc(&clk, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4); cudaMalloc(&dump, 128);
  const int iters = 20000;
  for (int threads = 128; threads <= 512; threads *= 2)
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) k<<<148, threads>>>(mode, iters, clk, sink, mode == 2 && rep == 0 ? dump : nullptr);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double c = (double)h[0] / iters;
    const int cells = mode == 0 ? 32 : 64, regs = mode == 1 ? 64 : 32;
    printf("%d warps, mode %d: %.1f clk per iteration; per SM: %.1f cell-bytes/clk, %.1f register-bytes/clk (%s)\n", threads / 32, mode, c,
           (double)threads * cells * 4 / c, (double)threads * regs * 4 / c,
           mode == 0 ? "x32: 32 cells -> 32 regs" : mode == 1 ? "2 x x32: 64 cells -> 64 regs" : "x32.pack::16b: 64 cells -> 32 regs");
  }
  uint32_t hd[32]; cudaMemcpy(hd, dump, sizeof(hd), cudaMemcpyDeviceToHost);
  printf("packed load, thread 5, regs 0..7: "); for (int i = 0; i < 8; ++i) printf("%08x ", hd[i]); printf("... reg 31: %08x\n", hd[31]);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
