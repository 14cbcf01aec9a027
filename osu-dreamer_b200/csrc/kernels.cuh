// Non-GEMM kernels of the denoiser path (HBM-bound elementwise / reduction work) and attention: host launchers.
#pragma once
#include "common.h"

namespace osd {

// latent model, inference half (latent.cu): fp32 channels-first building blocks
int launch_lat_block(const float* x, float* y, const float* const* w8, const float* film, int B, int L, cudaStream_t s);
// the same block with its two 1x1 convolutions on the tensor cores (latent_tc.cu): split-bf16 GEMMs + three streaming kernels
size_t lat_tc_pack_bytes();
size_t lat_tc_workspace_bytes(int B, int L);
int launch_lat_tc_pack(const float* w1, const float* b1, const float* w2, void* packed, cudaStream_t s);
int launch_lat_block_tc(const float* x, float* y, const float* const* w8, const void* packed, const float* film, void* ws,
                        int B, int L, cudaStream_t s);
int launch_lat_rmsnorm(const float* x, const float* gamma, float* y, int B, int C, long long N, int act, cudaStream_t s);
int launch_lat_conv1x1(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, long long N,
                       int act, int act_channels, cudaStream_t s);
int launch_lat_conv2d(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, int Ain, int L,
                      int kh, int sh, cudaStream_t s);
int launch_lat_down3(const float* x, const float* w, const float* bias, float* y, int B, int C, int L, cudaStream_t s);
int launch_lat_up3(const float* x, const float* w, const float* bias, float* y, int B, int C, int l, cudaStream_t s);
int launch_lat_mix(const float* x, const float* p, const float* g, float* y, int B, long long per_sample, int p_batch,
                   cudaStream_t s);
// style model inference (style.cu); params = 60 device pointers in the reference's state-dict order
size_t style_scratch_floats(int B);
int launch_style_forward(const float* const* params, const float* st, const float* labels, float* u, float* v, float* scratch,
                         int B, cudaStream_t s);
int launch_style_sample(const float* const* params, const float* labels, float* s_io, int num_steps, float* scratch,
                        float* eta_u0_out, int B, cudaStream_t s);
// style model training (style_train.cu): forward with saved activations, loss, backward (gradients accumulate into G[60])
size_t style_train_workspace_floats(int B);
int launch_style_train_forward(const float* const* params, const float* st, const float* labels, float* u, float* v, float* ws,
                               int B, cudaStream_t s);
int launch_style_loss(const float* st, const float* s1, const float* u, const float* v, float osl_w, float del_w, float* out4,
                      float* du, float* dv, float* acc_scratch, int B, cudaStream_t s);
int launch_style_backward(const float* const* params, const float* st, const float* labels, const float* du, const float* dv,
                          float* const* grads, float* ws, int B, cudaStream_t s);
size_t rope_table_floats(int L);  // [L][64] + the 32-row-transposed copy
int launch_rope_table(const float* inv_freq_host, int L, float* rope, cudaStream_t stream);
int launch_cf_to_tm(const float* in, void* out, int out_bf16, int B, int C, int L, cudaStream_t stream);
int launch_tm_to_cf(const void* in, int in_fp32, float* out, int B, int C, int L, cudaStream_t stream);
int launch_linear_small(const float* in, const float* W, const float* bias, float* out, int Bn, int N, int K,
                        int silu, cudaStream_t stream);
int launch_proj_in(const float* xt, const float* W, const float* bias, float* x, int B, int L, cudaStream_t stream);
int launch_prenorm_mod(const float* x, const float* mod, const void* cl, void* z, int z_fp32, int B, int L,
                       int cl_bcast, cudaStream_t stream, int split = 0, int cl_is_f32 = 0);
int launch_cf_to_tm_split(const float* in, void* out, int B, int C, int L, cudaStream_t stream);
int launch_pack_weight_split(const float* src, void* dst, int rows_src, int cols_src, int rows_dst, int cols_dst,
                             int split_at, int split_pad, cudaStream_t stream);
int launch_postnorm_gate_add(const float* x, const float* h, const float* mod, float* x_out, int B, int L,
                             cudaStream_t stream);
int launch_prenorm_mod_dwconv(const float* x, const float* mod, const float* wconv, const float* bconv, void* z,
                              int z_fp32, void* hmod_out, int B, int L, cudaStream_t stream, int split = 0);
int launch_swiglu_norm(const void* vg, void* hn, float* rinv_out, int is_fp32, int T, cudaStream_t stream);
int launch_final_norm_proj_out(const float* x, const float* Wo, const float* bo, float* v, int B, int L,
                               cudaStream_t stream);
size_t u_head_partial_floats(int B, int L);  // fpart: one 64-float row per u_head block
int launch_u_head(const float* xt, const float* const* w8, float* fpart, float* h1_save, float* h2pre_save, int B,
                  int L, cudaStream_t stream);
int launch_u_final(const float* fpart, float* fsum, const float* umod, const float* wout, const float* bout,
                   float u_scale, int L, float* u, int B, cudaStream_t stream);
int launch_sample_update(float* x, const float* v, const float* u, const float* eta_dev, int B, int L,
                         cudaStream_t stream);
int launch_sample_eta(const float* u, int B, float sqrt_c0, int num_steps, float* eta_out, cudaStream_t stream);
int launch_pack_weight(const float* src, void* dst, int dst_fp32, int rows_src, int cols_src, int rows_dst,
                       int cols_dst, int split_at, int split_pad, cudaStream_t stream);

// backward (elementwise_bwd.cu, uhead_bwd.cu)
int launch_final_bwd(const float* x, const float* dv, const float* Wo, float* dx, float* dWo, float* dbo, int B, int L,
                     cudaStream_t s);
int launch_postnorm_gate_bwd(const float* dx, const float* h, const float* mod, void* dh, float* dmod, float* dbias,
                             int B, int L, cudaStream_t s);
int launch_prenorm_mod_bwd(const void* dz, const float* x, const float* mod, float* dx, float* dmod, float* dbcl, int B,
                           int L, cudaStream_t s);
int launch_dwconv_prenorm_bwd(const void* dz2, const void* hmod, const float* x1, const float* mod, const float* wconv,
                              float* dx, float* dmod, float* dw, float* db, int B, int L, cudaStream_t s);
int launch_swiglu_norm_bwd(const void* vg, const void* dhn, const float* rinv, void* dvg, float* dbvg, int T,
                           cudaStream_t s);
int launch_qknorm_rope_bwd(void* dqkv, const void* raw, const float* rope, const float* qw, const float* kw,
                           float* dqw, float* dkw, float* dbias, const float* dq_acc, const int* dq_flag,
                           float dq_scale, int B, int L, cudaStream_t s);
int launch_proj_in_bwd(const float* dx, const float* xt, float* dW, float* db, int B, int L, cudaStream_t s);
int launch_silu_bwd(const float* da, const float* pre, void* dpre, float* db, int T, cudaStream_t s);
int launch_linear_small_bwd(const float* dout, const float* pre_or_null, const float* in, const float* W, float* dW,
                            float* db, float* din, float* dpre_scratch, int Bn, int N, int K, int silu,
                            cudaStream_t s);
int launch_unpack_grad(const float* src, float* dst, int rows_src, int cols_src, int rows_dst, int cols_dst,
                       int split_at, int split_pad, cudaStream_t s);
int launch_u_final_bwd(const float* du, const float* fsum, const float* umod, const float* wout, const float* bout,
                       float u_scale, int L, float* dfsum, float* dumod, float* dwout, float* dbout, int B,
                       cudaStream_t s);
int launch_u_head_bwd(const float* xt, const float* const* w8, const float* dfsum, float* const* g8, int B, int L,
                      cudaStream_t s);

// optimizer (optim.cu)
int launch_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, size_t n, int step, float lr,
                     float beta1, float beta2, float eps, float wd, float max_norm, float grad_scale, float ema_decay,
                     int ema_copy, double* acc_scratch, float* scal_out, cudaStream_t s);

// loss (loss.cu)
int launch_loss(const float* xt, const float* x1, const float* u, const float* v, float c0, float osl_w, float del_w,
                int B, int L, float* out4, float* du, float* dv, float* acc, cudaStream_t s);

// attention (attn_fwd.cu, attn_bwd.cu)
int launch_attn_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* dsum, void* dqkv, int B,
                    int L, int H, cudaStream_t stream);
// single-pass backward (attn_bwd_fused.cu): dq_acc fp32 [B*L, H*64] is scratch (zeroed and reduced into here),
// stats fp32 [attn_bwd_fused_stats_floats()] and dy_scaled bf16 [B*L, H*64] are scratch
size_t attn_bwd_fused_stats_floats(int B, int L, int H);
void attn_bwd_fused_set_trace(unsigned long long* buf, int cta);  // debugging aid, see osd_debug_attn_bwd_trace
// convert_dq = 0 leaves dq in dq_acc (fp32, unscaled; multiply by attn_bwd_fused_dq_scale()) unless the fallback flag
// (*attn_bwd_fused_flag(...) != 0) says the two-kernel path wrote dq into dqkv: launch_qknorm_rope_bwd consumes that.
int launch_attn_bwd_fused(const void* qkv, const void* y, const void* dy, const float* lse, float* stats, float* dq_acc,
                          void* dy_scaled, void* dqkv, int B, int L, int H, int convert_dq, cudaStream_t stream);
const int* attn_bwd_fused_flag(const float* stats, int B, int L, int H);
static inline float attn_bwd_fused_dq_scale() { return 0.125f; }
int launch_attn_fwd(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H, int variant,
                    cudaStream_t stream);
int launch_attn_fwd_db(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                       cudaStream_t stream);
int launch_attn_fwd_db_qt(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream);  // variant 7: Q tile resident in TMEM (A operand of S = Q K^T)
int launch_attn_fwd_db_dr(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream);  // variant 8: + pre-scaled Q (P = 2^S) and row sums by a ones-tile MMA
int launch_attn_fwd_db_pf(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream);  // variant 6: early barrier probes + S prefetch
// variant 15 (16: a quarter of the exponentials on the FMA pipe): ONE CTA per SM, two q tiles ping-pong against shared
// 128-row kv tiles, S / P decoupled in TMEM, SIXTEEN softmax warps (each score row split between two threads); fixed-bound
// softmax only, hands the online case to the "db" kernel on the device
int launch_attn_fwd_pp3(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H, int emu,
                        cudaStream_t stream);
// variant 17: variant 7 with the K / V tiles shared by a 2-CTA cluster (each CTA loads half a tile, TMA multicast)
int launch_attn_fwd_db_mc(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream);
// fp32-grade attention on the db pipeline: S = Q_hi K_hi^T + Q_lo K_hi^T, O += P_hi V_hi + P_hi V_lo (attn_fwd_db.cu, X3)
int launch_attn_fwd_db_x3(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                          cudaStream_t stream);
int launch_attn_fwd_x3(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H,
                       cudaStream_t stream);
int launch_qk_bound(const float* qw, const float* kw, float* out, cudaStream_t stream);

}  // namespace osd
