// Non-GEMM kernels of the denoiser path (HBM-bound elementwise / reduction work), host launchers.
#pragma once
#include "common.h"

namespace osd {

int launch_rope_table(const float* inv_freq_host, int L, float* rope, cudaStream_t stream);

}  // namespace osd
