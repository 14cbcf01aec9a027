// HBM-bound kernels: norms, adaLN modulation, depthwise conv, SwiGLU, layout changes, heads.
#include "kernels.cuh"
#include "ptx.cuh"

namespace osd {

struct InvFreq {
  float v[32];
};

// rope[l][0][i] = cos(l*f_i), rope[l][1][i] = sin(l*f_i); the product l*f_i is rounded to fp32 first,
// as torch.outer(arange(N).float(), inv_freq) does in osu_dreamer/common/attn.py:16-21.
__global__ void rope_table_kernel(InvFreq f, int L, float* __restrict__ rope) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * 32) return;
  const int l = idx >> 5, i = idx & 31;
  const float ang = __fmul_rn(static_cast<float>(l), f.v[i]);
  float s, c;
  sincosf(ang, &s, &c);
  rope[(size_t)l * 64 + i] = c;
  rope[(size_t)l * 64 + 32 + i] = s;
}

int launch_rope_table(const float* inv_freq_host, int L, float* rope, cudaStream_t stream) {
  OSD_CHECK(inv_freq_host && rope && L > 0, "rope_table: bad arguments");
  InvFreq f;
  for (int i = 0; i < 32; ++i) f.v[i] = inv_freq_host[i];
  const int n = L * 32;
  rope_table_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(f, L, rope);
  OSD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace osd
