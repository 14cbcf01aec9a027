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

/* ---- model dimensions (osu_dreamer/models/diffusion/model.yml:77-90).  The widths are compile-time constants of the
 * kernels; the backbone depth (BackboneArgs.depth, backbone.py:19) is a run-time property carried by `mode` ---- */
#define OSD_E 6        /* emb_dim       */
#define OSD_A 128      /* a_dim         */
#define OSD_S 32       /* style_dim     */
#define OSD_CG 512     /* global_cond   */
#define OSD_D 512      /* backbone_dim  */
#define OSD_H 16       /* n_heads       */
#define OSD_HD 64      /* head_dim      */
#define OSD_DH 1024    /* n_heads*head_dim */
#define OSD_DEPTH 8      /* default depth (model.yml:85): what a `mode` with depth bits 0 means */
#define OSD_MAX_DEPTH 32
#define OSD_HID 1365   /* int(512*4*2/3), osu_dreamer/common/swiglu.py:18 */
#define OSD_HIDP 1408  /* HID padded to a multiple of 64 (internal) */
#define OSD_U 64       /* u_head_dim    */

/* precision modes */
#define OSD_BF16 0     /* bf16 tensor-core operands, fp32 accumulate / residual / statistics */
#define OSD_F32X3 1    /* fp32-grade: every tensor-core product as the 3-term bf16 split a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
                          (operands stored as (hi | lo) bf16 pairs); inference only */

/* Every `mode` argument below: bits 0-7 = precision, bits 8-15 = backbone depth (0 selects OSD_DEPTH).  A model of depth
 * d has 20 + 18 * d parameter tensors (osd_num_params): the 6 leading tensors, 18 per layer, the 14 tail tensors. */
#define OSD_MODE(precision, depth) ((precision) | ((depth) << 8))

OSD_API int osd_abi_version(void);
OSD_API const char* osd_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
OSD_API unsigned long long osd_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Building blocks (exported for unit tests and micro-benchmarks; the model entry points below are
 * what the Python mirror of DiffusionModel calls).
 * ---------------------------------------------------------------------------------------------- */

/* C[M,N] (+)= A * B^T on tcgen05.  major 0: operand is [rows, K] row-major (K contiguous);
 * major 1: operand is [K, rows] row-major (rows contiguous).  elem: OSD_BF16 / OSD_TF32.
 * epi: 0 store(acc+bias), 1 silu(acc+bias), 2 atomic add into fp32 C (split_k >= 1).
 * bf16 problems with K >= 512 and enough [256 x 256] tiles run on CTA pairs (tcgen05 ... cta_group::2, M = 256 over a
 * 2-CTA cluster, half of the B tile staged per CTA); the environment variable OSD_GEMM_PAIR=0 / 1 switches that off / forces it
 * for every eligible shape (A/B measurements).
 * Replaces the reference's 1x1 Conv1d / Linear call sites. */
OSD_API int osd_gemm(const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb, void* C,
             int64_t ldc, int c_fp32, const float* bias, int M, int N, int K, int elem, int epi, int split_k,
             void* stream);
/* the split_k the library's own weight-gradient GEMMs use for an [M x N] reduce-add product over K: whole waves of the
 * persistent grid (CTAs, or CTA pairs = cta_group::2 [256 x 256] tiles where osd_gemm runs them: bf16, N % 256 == 0, more than
 * 8 k-blocks of 64 per work item) */
OSD_API int osd_gemm_split_k(int M, int N, int K);

/* qkv projection with the fused epilogue: bias + per-head RMSNorm(q), RMSNorm(k) + RoPE
 * (osu_dreamer/common/attn.py:75-81).  x [T,512], w [3072,512], out bf16 [T,3072];
 * rope = the table written by osd_rope_table (osd_rope_table_floats(L) floats: [L][2][32] cos|sin followed by a copy
 * transposed per 32 positions, which the epilogue reads); raw_out (optional) receives the pre-norm
 * projections for the backward pass. */
OSD_API int osd_qkv_proj(const void* x, const void* w, const float* bias, const float* qnorm_w, const float* knorm_w,
                 const float* rope, void* out, void* raw_out, int T, int L, int elem, void* stream);

/* rope[l][0][i] = cos(l * inv_freq[i]), rope[l][1][i] = sin(l * inv_freq[i]) with the angle formed in
 * fp32 exactly as osu_dreamer/common/attn.py:16-21 does; inv_freq_host are the 32 fp32 values
 * 10000 ** (arange(0,64,2)/-64). */
OSD_API int osd_rope_table(const float* inv_freq_host, int L, float* rope, void* stream);
/* size in floats of the buffer osd_rope_table fills (L*64 + ceil(L/32)*2048): every `rope` argument of this library
 * points at such a buffer */
OSD_API size_t osd_rope_table_floats(int L);

/* Bidirectional flash attention (head_dim 64, bf16): replaces F.scaled_dot_product_attention at
 * osu_dreamer/common/attn.py:82.  qkv bf16 [B*L, 3*H*64] token-major (q | k | v column blocks, head h at
 * columns h*64..h*64+63 of its block); y bf16 [B*L, H*64]; lse fp32 [B, H, L] (natural log, nullable).
 * bound_log2: optional DEVICE scalar, an upper bound of the scaled scores in log2 units (q,k are RMS-normalised
 * in this model, so such a bound exists per layer): selects the fixed-max softmax; NULL -> online softmax.
 * variant: 7 is the kernel the model runs (128 q rows per CTA, 64-row kv tiles, two S accumulators ping-pong, Q tile
 * resident in TMEM, 2 CTAs per SM; csrc/attn_fwd_db.cu).  Kept for A/B measurements: 4 = the same with Q in shared memory,
 * 6 = 4 + early barrier probes and S prefetch, 8 = 7 + pre-scaled Q and row sums by a ones-tile MMA (template siblings in the
 * same file); 15 = the cuDNN-shaped layout: one CTA per SM, two q tiles ping-pong against shared 128-row kv tiles, S / P
 * decoupled in TMEM, sixteen softmax warps (16 = the same with a quarter of the exponentials on the FMA pipe;
 * csrc/attn_fwd_pp3.cu) -- within 1 % of variant 7 in time because both sit at the board's power cap.
 * Anything else is an error. */
OSD_API int osd_attn_fwd(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                         int variant, void* stream);

/* Layout changes between the reference's channels-first [B, C, L] fp32 tensors and the internal
 * token-major [B*L, C] operands (bf16 when *_fp32 == 0). */
OSD_API int osd_tokens_to_channels(const void* in, int in_fp32, float* out, int B, int C, int L, void* stream);
OSD_API int osd_channels_to_tokens(const float* in, void* out, int out_fp32, int B, int C, int L, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model entry points.  `params` is a HOST array of 164 DEVICE pointers to the fp32 parameters in the
 * reference state-dict order (proj_audio.0.weight ... u_out.bias).
 * ---------------------------------------------------------------------------------------------- */
#define OSD_NUM_PARAMS 164

/* Sizes the caller must allocate (device memory, 1024-byte aligned). */
OSD_API int osd_num_params(int mode);                      /* 20 + 18 * depth (164 at the default depth) */
OSD_API size_t osd_packed_bytes(int mode);                 /* tensor-core operand copies of the weights */
OSD_API size_t osd_cond_floats(int B);                     /* cg + all adaLN vectors + u_mod (fp32), default depth */
OSD_API size_t osd_cond_floats_mode(int B, int mode);      /* the same for the depth carried by mode */
OSD_API size_t osd_workspace_bytes(int B, int L, int a_batch, int mode, int save);
OSD_API size_t osd_sample_extra_bytes(int B, int L, int a_batch);             /* default depth */
OSD_API size_t osd_sample_extra_bytes_mode(int B, int L, int a_batch, int mode);

/* Convert the fp32 parameters into padded tensor-core operands (call after every optimizer step /
 * load_state_dict).  HID 1365 is padded to 1408 with zero rows/columns (swiglu.py:18). */
OSD_API int osd_pack_weights(const float* const* params, void* packed, int mode, void* stream);

/* DiffusionModel._precompute_conditioning (model.py:73-84): audio [a_batch,128,L] fp32 channels-first,
 * style [B,32] -> a_tok [a_batch*L,128] operand dtype (token-major silu(proj_audio)), cond = cg [B,512] |
 * ssg1 of 8 layers [8][B,1536] | ssg2 [8][B,1536] | u_mod [B,128].  scratch >= a_batch*L*128*4 bytes. */
OSD_API int osd_precompute_conditioning(const float* const* params, const void* packed, int mode, const float* audio,
                                        int a_batch, const float* style, int B, int L, void* scratch, void* a_tok,
                                        float* cond, void* stream);

/* The (a, cg) arguments of DiffusionModel._pred (model.py:86-103) exactly as the reference passes them -- a
 * [a_batch,128,L] and cg [B,512], fp32 channels-first, possibly edited by the caller since _precompute_conditioning --
 * turned into what osd_pred_forward consumes: a_tok (token-major operand; (hi | lo) pairs in OSD_F32X3) and the cond pack
 * (cg copied to its head, then ssg1 / ssg2 of every layer and u_mod evaluated on it). */
OSD_API int osd_conditioning_from(const float* const* params, int mode, const float* a, int a_batch, const float* cg,
                                  int B, int L, void* a_tok, float* cond, void* stream);

/* DiffusionModel._pred (model.py:86-103): xt [B,6,L] fp32 -> u [B], v [B,6,L] fp32.  With save != 0 the
 * workspace (osd_workspace_bytes(..., save=1)) keeps every activation osd_pred_backward needs. */
OSD_API int osd_pred_forward(const float* const* params, const void* packed, int mode, const void* a_tok,
                             const float* cond, const float* rope, const float* xt, float* u, float* v, int B, int L,
                             int a_batch, void* workspace, int save, void* stream);

/* DiffusionModel.sample (model.py:117-138): x [B,6,L] holds the initial noise on entry and the sample on
 * exit; num_steps + 1 forwards, the u0 probe and eta stay on the device (eta_u0_out: 2 floats, nullable).
 * workspace = osd_workspace_bytes(..., save=0), extra = osd_sample_extra_bytes(). */
OSD_API int osd_sample(const float* const* params, const void* packed, int mode, const void* a_tok, const float* cond,
                       const float* rope, float* x, int num_steps, float c0, int B, int L, int a_batch,
                       void* workspace, void* extra, float* eta_u0_out, void* stream);

/* Backward of the attention kernel (gradient of attn.py:82): dqkv bf16 [B*L, 3*H*64] receives (dq | dk | dv)
 * w.r.t. the roped q, k and v; dsum fp32 [B,H,L] is scratch (rowsum(dO o O)). */
OSD_API int osd_attn_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* dsum, void* dqkv,
                         int B, int L, int H, void* stream);

/* Same gradient in one pass (five GEMMs, one exponential per score): dq is accumulated in fp32 through TMA
 * reduce-add into dq_acc [B*L, H*64] (scratch, zeroed by the call) and converted to bf16 into dqkv.
 * Scratch: stats fp32 [osd_attn_bwd_fused_stats_floats(B, L, H)], dy_scaled bf16 [B*L, H*64].
 * If the softmax statistics of some 128-row q tile span more than 96 octaves the call runs the two-kernel
 * path of osd_attn_bwd instead (decided on the device, no host synchronisation). */
OSD_API size_t osd_attn_bwd_fused_stats_floats(int B, int L, int H);
OSD_API int osd_attn_bwd_fused(const void* qkv, const void* y, const void* dy, const float* lse, float* stats,
                               float* dq_acc, void* dy_scaled, void* dqkv, int B, int L, int H, void* stream);

/* Debugging aid (tools/trace_attn_bwd.py): record the event timeline of CTA `cta` of the single-pass attention
 * backward into buf (DEVICE memory, 3 x 1024 u64 records: (event << 48) | (tile << 32) | SM clock); null = off. */
OSD_API void osd_debug_attn_bwd_trace(unsigned long long* buf, int cta);

/* ---- latent model, inference half: LatentModel.audio_encoder before and LatentModel.decode after diffusion.sample in
 * LDM.sample (osu_dreamer/models/inference/model.py:47,51; models/latent/model.py:53,103-133, unet.py, spec_features.py;
 * h_dim 128, expand 4 -> hidden 341, radius 2, stride 3 -- models/latent/model.yml:88-101).  fp32, channels-first
 * [B, C, L] like the reference; x and y are distinct caller-owned device buffers. */
/* one residual SwiGLU block of `layer` (unet.py:50-54): y = x + RMSNorm_g2(SwiGLU(RMSNorm_g1(x)(1+scale)+shift))(1+gate).
 * w8 = HOST array of 8 device pointers {norms.j.gamma, blocks.j.0.proj_vg.0.weight [128,1,5], .bias, proj_vg.1.weight
 * [682,128,1], .bias, proj_o.weight [128,341,1], .bias, blocks.j.1.gamma}; film = films.j(cond) [B,384] (scale | shift |
 * gate) or NULL for the unconditional layers. */
OSD_API int osd_lat_block(const float* x, float* y, const float* const* w8, const float* film, int B, int L, void* stream);
/* The same block with its two 1x1 convolutions (128 -> 682, 341 -> 128: 96 % of its FLOPs) as tcgen05 GEMMs on split-TF32
 * operands (3xTF32: x = hi + lo with hi = tf32(x), hi*hi + lo*hi + hi*lo accumulated in fp32, ~1e-6 of the fp32 result) and the norms /
 * FiLM / depthwise conv / SwiGLU as three streaming kernels; what LatentModel.audio_encoder / decode run by default
 * (osd_lat_block stays as the exact-fp32 CUDA-core version).  `packed` = osd_lat_tc_pack_bytes() bytes written once per
 * block by osd_lat_tc_pack from proj_vg.1.weight, proj_vg.1.bias and proj_o.weight; `ws` = osd_lat_tc_workspace_bytes(B, L)
 * bytes of scratch.  Arguments otherwise as osd_lat_block. */
OSD_API size_t osd_lat_tc_pack_bytes(void);
OSD_API size_t osd_lat_tc_workspace_bytes(int B, int L);
OSD_API int osd_lat_tc_pack(const float* w1, const float* b1, const float* w2, void* packed, void* stream);
OSD_API int osd_lat_block_tc(const float* x, float* y, const float* const* w8, const void* packed, const float* film, void* ws,
                             int B, int L, void* stream);
/* rms_norm over dim 1 (common/rms_norm.py:7-16) with optional gamma [C] and optional SiLU (act = 1); N = product of the
 * trailing dims */
OSD_API int osd_lat_rmsnorm(const float* x, const float* gamma, float* y, int B, int C, long long N, int act, void* stream);
/* pointwise Conv1d / Linear: y[b,o,n] = act(bias[o] + sum_i W[o,i] x[b,i,n]); act 0 none, 1 SiLU, 2 sigmoid on channels
 * < act_channels (LatentModel.decode's hit signals, model.py:127-130) */
OSD_API int osd_lat_conv1x1(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, long long N,
                            int act, int act_channels, void* stream);
/* SpecFeatures' strided Conv2d (spec_features.py:20,23): kernel (kh,3), stride (sh,1), padding (1,1) */
OSD_API int osd_lat_conv2d(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, int Ain, int L,
                           int kh, int sh, void* stream);
/* UNetEncoder down (unet.py:60-64): depthwise Conv1d k=3 pad 1 + AvgPool1d(3) -> [B,C,L/3];
 * UNetDecoder up (unet.py:81-85): nearest Upsample x3 + depthwise Conv1d k=3 pad 1 -> [B,C,3l] */
OSD_API int osd_lat_down3(const float* x, const float* w, const float* bias, float* y, int B, int C, int L, void* stream);
OSD_API int osd_lat_up3(const float* x, const float* w, const float* bias, float* y, int B, int C, int l, void* stream);
/* mixer (unet.py:126): y = x + p * g; p has batch p_batch (1 = the broadcast audio skips of predict, or B) */
OSD_API int osd_lat_mix(const float* x, const float* p, const float* g, float* y, int B, long long per_sample, int p_batch,
                        void* stream);

/* ---- style model inference: the sampler that runs right before diffusion.sample in LDM.sample
 * (osu_dreamer/models/inference/model.py:48; osu_dreamer/models/style/model.py:72-119, style_dim 32, h_dim 256, depth 8,
 * expand 4, label_features 128 -- models/style/model.yml:68-75).  params = HOST array of OSD_STYLE_NUM_PARAMS device
 * pointers (fp32) in the reference's state-dict order: cond_proj_w, cond_proj_b, null_labels, rff.W, rff.b,
 * proj_in.{weight,bias}, proj_out.0.weight, proj_out.1.{weight,bias}, u_out.{weight,bias}, films.i.{weight,bias} (i < 8),
 * blocks.i.0.{weight,bias}, blocks.i.3.{weight,bias} (i < 8).  scratch: osd_style_scratch_floats(B) floats. */
#define OSD_STYLE_NUM_PARAMS 60
OSD_API size_t osd_style_scratch_floats(int B);
/* StyleModel.forward (model.py:81-99): st [B,32], labels [B,5] (values < 0 select the null embedding) -> u [B], v [B,32] */
OSD_API int osd_style_forward(const float* const* params, const float* st, const float* labels, float* u, float* v,
                              float* scratch, int B, void* stream);
/* StyleModel.sample (model.py:101-119): s_inout [B,32] holds the initial noise on entry and the style codes on exit;
 * the probe forward, eta and all num_steps updates stay on the device; eta_u0_out (nullable) receives {eta, u0}. */
OSD_API int osd_style_sample(const float* const* params, const float* labels, float* s_inout, int num_steps,
                             float* scratch, float* eta_u0_out, int B, void* stream);

/* ---- style model training: `fit-style` (osu_dreamer/models/style/train.py:48-109, scripts/fit_style.py).  Exact fp32 on the
 * CUDA cores: 18 GFLOP per step at batch 512 is launch latency, not throughput.
 * osd_style_train_forward = StyleModel.forward (model.py:81-99) keeping its activations in `workspace`
 * (osd_style_train_workspace_floats(B) floats); osd_style_loss = the distance-marching loss of StyleTrainer.forward
 * (train.py:70-88): out4 = {loss, osl, del, u_mape} and the gradients du [B], dv [B,32] (acc_scratch: 4 floats);
 * osd_style_backward = what autograd computes for that forward: given du, dv it ACCUMULATES into grads[60] (HOST array of
 * DEVICE fp32 pointers, parameter shapes, reference state-dict order; entries 3 and 4 -- the Fourier-feature buffers --
 * are not touched and may be NULL).  labels < 0 select the null embedding (label dropping is the caller's draw). */
OSD_API size_t osd_style_train_workspace_floats(int B);
OSD_API int osd_style_train_forward(const float* const* params, const float* st, const float* labels, float* u, float* v,
                                    float* workspace, int B, void* stream);
OSD_API int osd_style_loss(const float* st, const float* s1, const float* u_pred, const float* v_pred, float osl_weight,
                           float del_weight, float* out4, float* du, float* dv, float* acc_scratch, int B, void* stream);
OSD_API int osd_style_backward(const float* const* params, const float* st, const float* labels, const float* du,
                               const float* dv, float* const* grads, float* workspace, int B, void* stream);

/* Gradient of DiffusionModel.forward (what autograd computes for the reference at train.py:84 + Lightning's
 * backward): given du [B] and dv [B,6,L], ACCUMULATES the parameter gradients into grads[164] (HOST array of
 * DEVICE fp32 pointers, parameter shapes).  `workspace` is the save=1 workspace the forward filled;
 * bwd_workspace = osd_backward_workspace_bytes().  audio / style / xt are the forward's inputs (data: no
 * gradient is produced for them).  Training shape only: a_batch == B. */
OSD_API size_t osd_backward_workspace_bytes(int B, int L, int a_batch);
OSD_API int osd_pred_backward(const float* const* params, const void* packed, int mode, const void* a_tok,
                              const float* cond, const float* rope, const float* audio, const float* style,
                              const float* xt, const float* du, const float* dv, float* const* grads, int B, int L,
                              int a_batch, void* workspace, void* bwd_workspace, void* stream);

/* Fused optimizer tail of fit-denoiser on FLAT fp32 buffers of n elements: global-norm clip
 * (gradient_clip_val, model.yml:39; coefficient min(1, max/(norm+1e-6)) as torch.nn.utils.clip_grad_norm_),
 * torch.optim.AdamW update (train.py:111; step is 1-based) and the EMA of the parameters
 * (get_ema_multi_avg_fn(.99), train.py:67,126; ema_copy != 0 on the first update; ema may be NULL).
 * grad_scale multiplies the raw gradients first (1/world_size after a summing allreduce).
 * acc_scratch: 1 double, scal_out: 2 floats {grad norm, applied multiplier}; no host synchronisation. */
OSD_API int osd_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, size_t n, int step, float lr,
                               float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                               float grad_scale, float ema_decay, int ema_copy, double* acc_scratch, float* scal_out,
                               void* stream);

/* Fused distance-marching loss of DiffusionTrainer.forward (train.py:86-108): given xt, x1 [B,6,L], u_pred [B],
 * v_pred [B,6,L] writes out4 = {loss, osl, del, u_mape} and the gradients of loss w.r.t. u_pred (du [B]) and
 * v_pred (dv [B,6,L]).  scratch: 4*B floats.  The random draws (t, x0) stay with the caller's generator. */
OSD_API int osd_loss_fwd_bwd(const float* xt, const float* x1, const float* u_pred, const float* v_pred, float c0,
                             float osl_weight, float del_weight, int B, int L, float* out4, float* du, float* dv,
                             float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSD_B200_H */
