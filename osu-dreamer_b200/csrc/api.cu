// extern "C" surface of libosd_b200.so (declared in include/osd_b200.h).
#include "../../include/osd_b200.h"

#include "common.h"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace osd;

extern "C" {

int osd_abi_version(void) { return OSD_ABI_VERSION; }
const char* osd_last_error(void) { return get_error(); }
unsigned long long osd_launch_count(void) { return launch_count(); }

int osd_gemm(const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb, void* C,
             int64_t ldc, int c_fp32, const float* bias, int M, int N, int K, int elem, int epi, int split_k,
             void* stream) {
  OSD_CHECK(epi >= 0 && epi <= 2, "osd_gemm: epi %d not in {0,1,2}", epi);
  GemmArgs g;
  g.A = A; g.B = B; g.a_major = a_major; g.b_major = b_major; g.lda = lda; g.ldb = ldb;
  g.M = M; g.N = N; g.K = K; g.elem = elem; g.epi = epi; g.C = C; g.ldc = ldc; g.c_fp32 = c_fp32;
  g.bias = bias; g.split_k = split_k;
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int osd_gemm_split_k(int M, int N, int K) { return gemm_split_for(M, N, K); }

int osd_qkv_proj(const void* x, const void* w, const float* bias, const float* qnorm_w, const float* knorm_w,
                 const float* rope, void* out, void* raw_out, int T, int L, int elem, void* stream) {
  GemmArgs g;
  g.A = x; g.B = w; g.lda = OSD_D; g.ldb = OSD_D; g.M = T; g.N = 3 * OSD_DH; g.K = OSD_D; g.elem = elem;
  g.epi = EPI_QKV; g.C = out; g.ldc = 3 * OSD_DH; g.c_fp32 = 0; g.bias = bias; g.qnorm_w = qnorm_w;
  g.knorm_w = knorm_w; g.rope = rope; g.L = L; g.dh = OSD_DH; g.raw_out = raw_out;
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

size_t osd_rope_table_floats(int L) { return rope_table_floats(L); }

int osd_lat_block(const float* x, float* y, const float* const* w8, const float* film, int B, int L, void* stream) {
  return launch_lat_block(x, y, w8, film, B, L, static_cast<cudaStream_t>(stream));
}
size_t osd_lat_tc_pack_bytes(void) { return lat_tc_pack_bytes(); }
size_t osd_lat_tc_workspace_bytes(int B, int L) { return lat_tc_workspace_bytes(B, L); }
int osd_lat_tc_pack(const float* w1, const float* b1, const float* w2, void* packed, void* stream) {
  return launch_lat_tc_pack(w1, b1, w2, packed, static_cast<cudaStream_t>(stream));
}
int osd_lat_block_tc(const float* x, float* y, const float* const* w8, const void* packed, const float* film, void* ws, int B,
                     int L, void* stream) {
  return launch_lat_block_tc(x, y, w8, packed, film, ws, B, L, static_cast<cudaStream_t>(stream));
}
int osd_lat_rmsnorm(const float* x, const float* gamma, float* y, int B, int C, long long N, int act, void* stream) {
  return launch_lat_rmsnorm(x, gamma, y, B, C, N, act, static_cast<cudaStream_t>(stream));
}
int osd_lat_conv1x1(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, long long N, int act,
                    int act_channels, void* stream) {
  return launch_lat_conv1x1(x, W, bias, y, B, Cin, Cout, N, act, act_channels, static_cast<cudaStream_t>(stream));
}
int osd_lat_conv2d(const float* x, const float* W, const float* bias, float* y, int B, int Cin, int Cout, int Ain, int L, int kh,
                   int sh, void* stream) {
  return launch_lat_conv2d(x, W, bias, y, B, Cin, Cout, Ain, L, kh, sh, static_cast<cudaStream_t>(stream));
}
int osd_lat_down3(const float* x, const float* w, const float* bias, float* y, int B, int C, int L, void* stream) {
  return launch_lat_down3(x, w, bias, y, B, C, L, static_cast<cudaStream_t>(stream));
}
int osd_lat_up3(const float* x, const float* w, const float* bias, float* y, int B, int C, int l, void* stream) {
  return launch_lat_up3(x, w, bias, y, B, C, l, static_cast<cudaStream_t>(stream));
}
int osd_lat_mix(const float* x, const float* p, const float* g, float* y, int B, long long per_sample, int p_batch, void* stream) {
  return launch_lat_mix(x, p, g, y, B, per_sample, p_batch, static_cast<cudaStream_t>(stream));
}

size_t osd_style_scratch_floats(int B) { return style_scratch_floats(B); }
int osd_style_forward(const float* const* params, const float* st, const float* labels, float* u, float* v, float* scratch,
                      int B, void* stream) {
  return launch_style_forward(params, st, labels, u, v, scratch, B, static_cast<cudaStream_t>(stream));
}
int osd_style_sample(const float* const* params, const float* labels, float* s_inout, int num_steps, float* scratch,
                     float* eta_u0_out, int B, void* stream) {
  return launch_style_sample(params, labels, s_inout, num_steps, scratch, eta_u0_out, B, static_cast<cudaStream_t>(stream));
}

size_t osd_style_train_workspace_floats(int B) { return style_train_workspace_floats(B); }
int osd_style_train_forward(const float* const* params, const float* st, const float* labels, float* u, float* v, float* workspace,
                            int B, void* stream) {
  return launch_style_train_forward(params, st, labels, u, v, workspace, B, static_cast<cudaStream_t>(stream));
}
int osd_style_loss(const float* st, const float* s1, const float* u_pred, const float* v_pred, float osl_weight, float del_weight,
                   float* out4, float* du, float* dv, float* acc_scratch, int B, void* stream) {
  return launch_style_loss(st, s1, u_pred, v_pred, osl_weight, del_weight, out4, du, dv, acc_scratch, B,
                           static_cast<cudaStream_t>(stream));
}
int osd_style_backward(const float* const* params, const float* st, const float* labels, const float* du, const float* dv,
                       float* const* grads, float* workspace, int B, void* stream) {
  return launch_style_backward(params, st, labels, du, dv, grads, workspace, B, static_cast<cudaStream_t>(stream));
}

int osd_rope_table(const float* inv_freq_host, int L, float* rope, void* stream) {
  return launch_rope_table(inv_freq_host, L, rope, static_cast<cudaStream_t>(stream));
}

int osd_attn_fwd(const void* qkv, void* y, float* lse, const float* bound_log2, int B, int L, int H, int variant,
                 void* stream) {
  return launch_attn_fwd(qkv, y, lse, bound_log2, B, L, H, variant, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
