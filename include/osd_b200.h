/* osd_b200.h -- C ABI of the B200-native osu-dreamer denoiser hot path (libosd_b200.so).
 *
 * The reference (jaswon/osu-dreamer @ 391c4312) has no FFI: its boundary for this path is the Python
 * class osu_dreamer/models/diffusion/model.py:23 `DiffusionModel` (+ `DiffusionTrainer`,
 * osu_dreamer/models/diffusion/train.py:33).  These entry points are what a ctypes binding of that class
 * calls; each one cites the reference code it replaces.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless named *_host.
 *  - all memory (inputs, outputs, weights, saved activations, workspace) is owned by the caller; the
 *    library never allocates, frees or retains device memory.
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*); no internal synchronisation.
 *  - return 0 on success; non-zero on error with a thread-local message in osd_last_error().
 *  - there is NO CPU fallback: every entry point requires an sm_100 device.
 *  - internal activation layout is token-major [B*L, C] (C contiguous); the public tensors keep the
 *    reference's channels-first [B, C, L] layout.
 */
#ifndef OSD_B200_H
#define OSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSD_ABI_VERSION 1

#if defined(__GNUC__)
#define OSD_API __attribute__((visibility("default")))
#else
#define OSD_API
#endif

/* ---- model dimensions (osu_dreamer/models/diffusion/model.yml:77-90; fixed at compile time) ---- */
#define OSD_E 6        /* emb_dim       */
#define OSD_A 128      /* a_dim         */
#define OSD_S 32       /* style_dim     */
#define OSD_CG 512     /* global_cond   */
#define OSD_D 512      /* backbone_dim  */
#define OSD_H 16       /* n_heads       */
#define OSD_HD 64      /* head_dim      */
#define OSD_DH 1024    /* n_heads*head_dim */
#define OSD_DEPTH 8
#define OSD_HID 1365   /* int(512*4*2/3), osu_dreamer/common/swiglu.py:18 */
#define OSD_HIDP 1408  /* HID padded to a multiple of 64 (internal) */
#define OSD_U 64       /* u_head_dim    */

/* precision modes */
#define OSD_BF16 0     /* bf16 tensor-core operands, fp32 accumulate / residual / statistics */
#define OSD_TF32 1     /* fp32 storage, tf32 tensor-core operands */

OSD_API int osd_abi_version(void);
OSD_API const char* osd_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Building blocks (exported for unit tests and micro-benchmarks; the model entry points below are
 * what the Python mirror of DiffusionModel calls).
 * ---------------------------------------------------------------------------------------------- */

/* C[M,N] (+)= A * B^T on tcgen05.  major 0: operand is [rows, K] row-major (K contiguous);
 * major 1: operand is [K, rows] row-major (rows contiguous).  elem: OSD_BF16 / OSD_TF32.
 * epi: 0 store(acc+bias), 1 silu(acc+bias), 2 atomic add into fp32 C (split_k >= 1).
 * Replaces the reference's 1x1 Conv1d / Linear call sites. */
OSD_API int osd_gemm(const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb, void* C,
             int64_t ldc, int c_fp32, const float* bias, int M, int N, int K, int elem, int epi, int split_k,
             void* stream);

/* qkv projection with the fused epilogue: bias + per-head RMSNorm(q), RMSNorm(k) + RoPE
 * (osu_dreamer/common/attn.py:75-81).  x [T,512], w [3072,512], out bf16 [T,3072];
 * rope table [L][2][32] fp32 (cos|sin) from osd_rope_table; raw_out (optional) receives the pre-norm
 * projections for the backward pass. */
OSD_API int osd_qkv_proj(const void* x, const void* w, const float* bias, const float* qnorm_w, const float* knorm_w,
                 const float* rope, void* out, void* raw_out, int T, int L, int elem, void* stream);

/* rope[l][0][i] = cos(l * inv_freq[i]), rope[l][1][i] = sin(l * inv_freq[i]) with the angle formed in
 * fp32 exactly as osu_dreamer/common/attn.py:16-21 does; inv_freq_host are the 32 fp32 values
 * 10000 ** (arange(0,64,2)/-64). */
OSD_API int osd_rope_table(const float* inv_freq_host, int L, float* rope, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSD_B200_H */
