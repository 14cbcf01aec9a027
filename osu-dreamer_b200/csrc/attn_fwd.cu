// Forward attention dispatch.  The kernels live in attn_fwd_db.cu (the model's: variant 7, and its A/B template siblings 4 / 6 /
// 8) and attn_fwd_pp3.cu (the cuDNN-shaped layout: 15 / 16).  The first-generation single-accumulator kernels (variants 0-3:
// 64- / 128-row kv tiles, P staged in shared memory or kept in TMEM) were removed in round 2 once the design had settled;
// their measurements are in DESIGN.md 5.
#include "kernels.cuh"

namespace osd {

int launch_attn_fwd(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H, int variant,
                    cudaStream_t stream) {
  OSD_CHECK(qkv && y && B > 0 && L > 0 && H > 0, "attn_fwd: bad arguments");
  if (variant == 7) return launch_attn_fwd_db_qt(qkv, y, lse, bound_log2, B, L, H, stream);        // Q in TMEM: the model's kernel
  if (variant == 17) return launch_attn_fwd_db_mc(qkv, y, lse, bound_log2, B, L, H, stream);        // 7 + K/V multicast in a CTA pair
  // 2 q tiles / CTA, 16 softmax warps; 15: all exponentials on the SFU, 16 / 18 / 19 / 20: 2 / 1 / 3 / 4 eighths of them on the FMA pipe
  if (variant == 15) return launch_attn_fwd_pp3(qkv, y, lse, bound_log2, B, L, H, 0, stream);
  if (variant == 16) return launch_attn_fwd_pp3(qkv, y, lse, bound_log2, B, L, H, 2, stream);
  if (variant == 18) return launch_attn_fwd_pp3(qkv, y, lse, bound_log2, B, L, H, 1, stream);
  if (variant == 19) return launch_attn_fwd_pp3(qkv, y, lse, bound_log2, B, L, H, 3, stream);
  if (variant == 20) return launch_attn_fwd_pp3(qkv, y, lse, bound_log2, B, L, H, 4, stream);
  if (variant == 4) return launch_attn_fwd_db(qkv, y, lse, bound_log2, B, L, H, stream);           // double-buffered S, Q in smem
  if (variant == 6) return launch_attn_fwd_db_pf(qkv, y, lse, bound_log2, B, L, H, stream);        // + probes / S prefetch
  if (variant == 8) return launch_attn_fwd_db_dr(qkv, y, lse, bound_log2, B, L, H, stream);        // + direct exponent
  OSD_CHECK(false, "attn_fwd: unknown variant %d (4, 6, 7, 8, 15-20)", variant);
  return 1;
}

}  // namespace osd
