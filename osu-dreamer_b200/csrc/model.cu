// Denoiser forward orchestration (DiffusionModel._precompute_conditioning / _pred / sample of the reference,
// osu_dreamer/models/diffusion/model.py:73-138) over the kernels in gemm.cu / attn_fwd.cu / elementwise.cu.
#include "model.cuh"

#include "gemm.cuh"
#include "kernels.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace osd {

static inline size_t esz_of(int) { return 2; }              // operands are bf16 in both modes
static inline int kmul_of(int mode) { return mode == OSD_F32X3 ? 2 : 1; }  // (hi | lo) pairs double the K extent
static inline size_t al(size_t x) { return align_up(x, 1024); }

PackedLayout packed_layout(int mode, int depth) {
  PackedLayout p;
  p.esz = esz_of(mode);
  p.kmul = kmul_of(mode);
  const size_t e = p.esz * p.kmul;
  size_t off = 0;
  p.wa = off;
  off += al((size_t)128 * 128 * e);
  size_t lo = 0;
  p.l_cl = lo;
  lo += al((size_t)512 * 128 * e);
  p.l_qkv = lo;
  lo += al((size_t)3072 * 512 * e);
  p.l_out = lo;
  lo += al((size_t)512 * 1024 * e);
  p.l_vg = lo;
  lo += al((size_t)2 * OSD_HIDP * 512 * e);
  p.l_po = lo;
  lo += al((size_t)512 * OSD_HIDP * e);
  p.layer0 = off;
  p.layer_stride = lo;
  off += depth * lo;
  p.bvg = off;
  off += al((size_t)depth * 2 * OSD_HIDP * 4);
  p.bounds = off;
  off += al(depth * 4);
  p.total = off;
  return p;
}

ActPlan make_plan(int B, int L, int a_batch, int mode, int save, int depth) {
  ActPlan a;
  const size_t T = (size_t)B * L, Ta = (size_t)a_batch * L;
  const bool X = mode == OSD_F32X3;
  const size_t e = esz_of(mode) * kmul_of(mode);  // bytes per operand element ((hi | lo) pairs in split mode)
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += al(bytes);
    return r;
  };
  a.x0 = take(T * 512 * 4);
  a.cl = take(Ta * 512 * (X ? 4 : 2));  // proj_cl output: bf16, or fp32 in split mode (added inside prenorm_mod)
  a.z = take(T * 512 * e);
  a.qkv_raw = save ? take(T * 3072 * 2) : 0;
  a.qkv = take(T * 3072 * e);
  a.y = take(T * 1024 * e);
  a.lse = take(T * 16 * 4);
  a.o = take(T * 512 * 4);
  a.x1 = take(T * 512 * 4);
  a.hmod = save ? take(T * 512 * 2) : 0;
  a.z2 = take(T * 512 * e);
  a.vg = take(T * 2 * OSD_HIDP * (X ? 4 : 2));  // consumed by swiglu_norm only: fp32 in split mode
  a.hn = take(T * OSD_HIDP * e);
  a.rinv2 = take(T * 4);
  a.f = take(T * 512 * 4);
  a.layer_bytes = o;
  a.layer_stride = save ? o : 0;
  size_t tot = save ? depth * o : o;
  auto take2 = [&](size_t bytes) {
    size_t r = tot;
    tot += al(bytes);
    return r;
  };
  a.x_final = save ? take2(T * 512 * 4) : a.x0;
  a.fsum = take2((size_t)B * 64 * 4);
  a.fpart = take2(u_head_partial_floats(B, L) * 4);
  a.uh1 = save ? take2(T * 64 * 4) : 0;
  a.uh2 = save ? take2(T * 64 * 4) : 0;
  a.total = tot;
  return a;
}

// ------------------------------------------------------------------------------------------------
static int pack_weights(const float* const* P, uint8_t* packed, int mode, int depth, cudaStream_t s) {
  const PackedLayout lay = packed_layout(mode, depth);
  const bool X = mode == OSD_F32X3;
  auto pack = [&](const float* src, void* dst, int rs, int cs, int rd, int cd, int sa, int sp) -> int {
    if (X) return launch_pack_weight_split(src, dst, rs, cs, rd, cd, sa, sp, s);
    return launch_pack_weight(src, dst, 0, rs, cs, rd, cd, sa, sp, s);
  };
  OSD_TRY(pack(P[P_AUDIO_W], packed + lay.wa, 128, 128, 128, 128, 0, 0));
  for (int l = 0; l < depth; ++l) {
    uint8_t* lb = packed + lay.layer0 + l * lay.layer_stride;
    OSD_TRY(pack(P[lp(l, L_CL_W)], lb + lay.l_cl, 512, 128, 512, 128, 0, 0));
    OSD_TRY(pack(P[lp(l, L_QKV_W)], lb + lay.l_qkv, 3072, 512, 3072, 512, 0, 0));
    OSD_TRY(pack(P[lp(l, L_OUT_W)], lb + lay.l_out, 512, 1024, 512, 1024, 0, 0));
    OSD_TRY(pack(P[lp(l, L_VG_W)], lb + lay.l_vg, 2 * OSD_HID, 512, 2 * OSD_HIDP, 512, OSD_HID, OSD_HIDP));
    OSD_TRY(pack(P[lp(l, L_PO_W)], lb + lay.l_po, 512, OSD_HID, 512, OSD_HIDP, 0, 0));
    OSD_TRY(launch_pack_weight(P[lp(l, L_VG_B)], packed + lay.bvg + (size_t)l * 2 * OSD_HIDP * 4, 1, 2 * OSD_HID, 1,
                               2 * OSD_HIDP, 1, OSD_HID, OSD_HIDP, s));
    OSD_TRY(launch_qk_bound(P[lp(l, L_QN_W)], P[lp(l, L_KN_W)], reinterpret_cast<float*>(packed + lay.bounds) + l, s));
  }
  return 0;
}

// modulation vectors of every layer (backbone.py:76,82) and u_mod (model.py:100) from cg = cond[0 : B*512]
static int modulation_from_cg(const float* const* P, int depth, float* cond, int B, cudaStream_t s) {
  CondPack c{cond, B, depth};
  for (int l = 0; l < depth; ++l) {
    OSD_TRY(launch_linear_small(c.cg(), P[lp(l, L_SSG1_W)], P[lp(l, L_SSG1_B)], const_cast<float*>(c.mod1(l)), B, 1536,
                                512, 0, s));
    OSD_TRY(launch_linear_small(c.cg(), P[lp(l, L_SSG2_W)], P[lp(l, L_SSG2_B)], const_cast<float*>(c.mod2(l)), B, 1536,
                                512, 0, s));
  }
  const int ts = tail_shift(depth);
  OSD_TRY(launch_linear_small(c.cg(), P[P_UMOD_W + ts], P[P_UMOD_B + ts], const_cast<float*>(c.umod()), B, 128, 512, 0, s));
  return 0;
}

static int precompute_conditioning(const float* const* P, const PackedW& W, int mode, int depth, const float* audio,
                                   int a_batch, const float* style, int B, int L, void* a_tok, float* cond, void* scratch,
                                   cudaStream_t s) {
  const int Ta = a_batch * L;
  const int X = mode == OSD_F32X3;
  const int km = X ? 2 : 1;
  // audio [Ba,128,L] fp32 -> token-major bf16 operand ((hi | lo) in split mode) -> a = silu(proj_audio(audio))
  // (model.py:45,82)
  if (X)
    OSD_TRY(launch_cf_to_tm_split(audio, scratch, a_batch, 128, L, s));
  else
    OSD_TRY(launch_cf_to_tm(audio, scratch, 1, a_batch, 128, L, s));
  GemmArgs g;
  g.A = scratch; g.B = W.wa(); g.lda = 128 * km; g.ldb = 128 * km; g.M = Ta; g.N = 128; g.K = 128;
  g.elem = ELEM_BF16; g.epi = EPI_SILU; g.C = a_tok; g.ldc = 128 * km; g.c_fp32 = 0; g.split3 = X; g.c_split = X;
  g.bias = P[P_AUDIO_B];
  OSD_TRY(launch_gemm(g, s));
  // cg = silu(proj_style(style)) (model.py:46,83)
  OSD_TRY(launch_linear_small(style, P[P_STYLE_W], P[P_STYLE_B], cond, B, 512, 32, 1, s));
  return modulation_from_cg(P, depth, cond, B, s);
}

// proj_cl for one layer: cl = a_tok * Wcl^T + b  -> bf16 [Ta, 512]   (backbone.py:63,78)
static int proj_cl(const float* const* P, const PackedW& W, int mode, int l, const void* a_tok, int Ta, void* cl,
                   cudaStream_t s) {
  const int X = mode == OSD_F32X3, km = X ? 2 : 1;
  GemmArgs g;
  g.A = a_tok; g.B = W.cl(l); g.lda = 128 * km; g.ldb = 128 * km; g.M = Ta; g.N = 512; g.K = 128;
  g.elem = ELEM_BF16; g.epi = EPI_STORE; g.C = cl; g.ldc = 512; g.c_fp32 = X; g.split3 = X;
  g.bias = P[lp(l, L_CL_B)];
  return launch_gemm(g, s);
}

// forward attention kernel of the model; OSD_ATTN_FWD=<n> forces one variant for A/B measurements.  Long sequences run the
// one-CTA-per-SM kernel with two q tiles and sixteen softmax warps (variant 18, attn_fwd_pp3.cu: -4 % kernel time at L = 8192,
// +3.7 % sampling throughput, profiles/r02x_*); shorter ones the two-CTA-per-SM kernel (variant 7, attn_fwd_db.cu), whose
// smaller CTAs balance better when there are few of them (tools/attn_fwd_threshold.py at 131 k tokens: variant 18 / 7 = 1.09 at
// L = 1024, 1.04 at 2048, 0.99 at 3072-4096, 0.97 from 5120 up).
static int attn_fwd_variant(int L) {
  static const int forced = [] {
    const char* e = getenv("OSD_ATTN_FWD");
    return (e != nullptr && e[0] >= '0' && e[0] <= '9') ? atoi(e) : -1;
  }();
  if (forced >= 0) return forced;
  return L >= 5120 ? 18 : 7;
}

// OSD_X3_KERNEL=old selects the single-buffered fp32-grade attention kernel (attn_fwd_x3.cu, selectable terms) for A/B
static bool x3_old_kernel() {
  static const bool v = [] {
    const char* e = getenv("OSD_X3_KERNEL");
    return e != nullptr && strcmp(e, "old") == 0;
  }();
  return v;
}

struct FwdCtx {
  const float* const* P;
  PackedW W;
  int mode, depth, B, L, a_batch;
  CondPack cond;
  const void* a_tok;  // [Ta,128] operand dtype
  const float* rope;
  uint8_t* ws;
  ActPlan plan;
  int save;
  const uint8_t* cl_hoisted;  // [depth][Ta,512] bf16 or null
};

static int pred_forward(const FwdCtx& c, const float* xt, float* u, float* v, cudaStream_t s) {
  const int B = c.B, L = c.L, T = B * L, Ta = c.a_batch * L;
  const int mode = c.mode;
  const int X = mode == OSD_F32X3;  // fp32-grade: split-precision operands (hi | lo), inference only
  const int km = X ? 2 : 1;
  OSD_CHECK(!(X && c.save), "pred_forward: the fp32-grade (3x bf16) path is inference-only");
  const ActPlan& pl = c.plan;
  auto LB = [&](int l) { return c.ws + (size_t)l * pl.layer_stride; };
  const void* a_tok = c.a_tok;
  const size_t cl_bytes = al((size_t)Ta * 512 * (X ? 4 : 2));

  const int depth = c.depth, ts = tail_shift(depth);
  OSD_TRY(launch_proj_in(xt, c.P[P_IN_W], c.P[P_IN_B], reinterpret_cast<float*>(LB(0) + pl.x0), B, L, s));
  for (int l = 0; l < depth; ++l) {
    uint8_t* lb = LB(l);
    float* x0 = reinterpret_cast<float*>(lb + pl.x0);
    float* x1 = reinterpret_cast<float*>(lb + pl.x1);
    float* xo = (l == depth - 1) ? reinterpret_cast<float*>(c.ws + pl.x_final)
                         : reinterpret_cast<float*>(LB(l + 1) + pl.x0);
    const void* cl;
    if (c.cl_hoisted != nullptr) {
      cl = c.cl_hoisted + (size_t)l * cl_bytes;
    } else {
      OSD_TRY(proj_cl(c.P, c.W, mode, l, a_tok, Ta, lb + pl.cl, s));
      cl = lb + pl.cl;
    }
    // ---- attention sub-block (backbone.py:76-80)
    OSD_TRY(launch_prenorm_mod(x0, c.cond.mod1(l), cl, lb + pl.z, 0, B, L, c.a_batch == 1, s, X, X));
    GemmArgs q;
    q.A = lb + pl.z; q.B = c.W.qkv(l); q.lda = 512 * km; q.ldb = 512 * km; q.M = T; q.N = 3072; q.K = 512;
    q.elem = ELEM_BF16; q.split3 = X; q.c_split = X;
    q.epi = EPI_QKV; q.C = lb + pl.qkv; q.ldc = 3072 * km; q.c_fp32 = 0; q.bias = c.P[lp(l, L_QKV_B)];
    q.qnorm_w = c.P[lp(l, L_QN_W)]; q.knorm_w = c.P[lp(l, L_KN_W)]; q.rope = c.rope; q.L = L; q.dh = 1024;
    q.raw_out = c.save ? lb + pl.qkv_raw : nullptr;
    OSD_TRY(launch_gemm(q, s));
    if (X)
      OSD_TRY((x3_old_kernel() ? launch_attn_fwd_x3 : launch_attn_fwd_db_x3)(lb + pl.qkv, lb + pl.y, nullptr, c.W.bound(l), B, L,
                                                                             16, s));
    else
      OSD_TRY(launch_attn_fwd(lb + pl.qkv, lb + pl.y, reinterpret_cast<float*>(lb + pl.lse), c.W.bound(l), B, L, 16,
                              attn_fwd_variant(L),
                              s));
    GemmArgs o;
    o.A = lb + pl.y; o.B = c.W.out(l); o.lda = 1024 * km; o.ldb = 1024 * km; o.M = T; o.N = 512; o.K = 1024;
    o.elem = ELEM_BF16; o.split3 = X;
    o.epi = EPI_STORE; o.C = lb + pl.o; o.ldc = 512; o.c_fp32 = 1; o.bias = c.P[lp(l, L_OUT_B)];
    OSD_TRY(launch_gemm(o, s));
    OSD_TRY(launch_postnorm_gate_add(x0, reinterpret_cast<float*>(lb + pl.o), c.cond.mod1(l), x1, B, L, s));
    // ---- ffn sub-block (backbone.py:82-86, swiglu.py:27-32)
    OSD_TRY(launch_prenorm_mod_dwconv(x1, c.cond.mod2(l), c.P[lp(l, L_DW_W)], c.P[lp(l, L_DW_B)], lb + pl.z2, 0,
                                      c.save ? lb + pl.hmod : nullptr, B, L, s, X));
    GemmArgs g;
    g.A = lb + pl.z2; g.B = c.W.vg(l); g.lda = 512 * km; g.ldb = 512 * km; g.M = T; g.N = 2 * OSD_HIDP; g.K = 512;
    g.elem = ELEM_BF16; g.split3 = X;
    g.epi = EPI_STORE; g.C = lb + pl.vg; g.ldc = 2 * OSD_HIDP; g.c_fp32 = X; g.bias = c.W.bvg(l);
    OSD_TRY(launch_gemm(g, s));
    OSD_TRY(launch_swiglu_norm(lb + pl.vg, lb + pl.hn, reinterpret_cast<float*>(lb + pl.rinv2), X ? 2 : 0, T, s));
    GemmArgs po;
    po.A = lb + pl.hn; po.B = c.W.po(l); po.lda = OSD_HIDP * km; po.ldb = OSD_HIDP * km; po.M = T; po.N = 512;
    po.K = OSD_HIDP; po.elem = ELEM_BF16; po.split3 = X;
    po.epi = EPI_STORE; po.C = lb + pl.f; po.ldc = 512; po.c_fp32 = 1; po.bias = c.P[lp(l, L_PO_B)];
    OSD_TRY(launch_gemm(po, s));
    OSD_TRY(launch_postnorm_gate_add(x1, reinterpret_cast<float*>(lb + pl.f), c.cond.mod2(l), xo, B, L, s));
  }
  OSD_TRY(launch_final_norm_proj_out(reinterpret_cast<float*>(c.ws + pl.x_final), c.P[P_OUT_W + ts], c.P[P_OUT_B + ts], v,
                                     B, L, s));
  // ---- distance head (model.py:99-102)
  const float* uw[8] = {c.P[P_UH0_W + ts], c.P[P_UH0_B + ts], c.P[P_UH1_W + ts], c.P[P_UH1_B + ts],
                        c.P[P_UH3_W + ts], c.P[P_UH3_B + ts], c.P[P_UH4_W + ts], c.P[P_UH4_B + ts]};
  float* fsum = reinterpret_cast<float*>(c.ws + pl.fsum);
  float* fpart = reinterpret_cast<float*>(c.ws + pl.fpart);
  OSD_TRY(launch_u_head(xt, uw, fpart, c.save ? reinterpret_cast<float*>(c.ws + pl.uh1) : nullptr,
                        c.save ? reinterpret_cast<float*>(c.ws + pl.uh2) : nullptr, B, L, s));
  OSD_TRY(launch_u_final(fpart, fsum, c.cond.umod(), c.P[P_UOUT_W + ts], c.P[P_UOUT_B + ts], sqrtf(2.0f * OSD_E), L, u, B,
                         s));
  return 0;
}

// ================================================================================================ backward
struct BwdPlan {
  size_t dx, dh, dhn, dvg, dz2, dy, dys, dqkv, dq_acc, dz, dsum, da_tok, a_pre, da_pre, audio_tm, dcond, dfsum, dpre_s;
  size_t gWvg, gWpo, gbvg;  // padded fp32 gradient scratch
  size_t total;
};
static BwdPlan make_bwd_plan(int B, int L, int a_batch, int depth) {
  BwdPlan p;
  const size_t T = (size_t)B * L, Ta = (size_t)a_batch * L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += al(bytes);
    return r;
  };
  p.dx = take(T * 512 * 4);
  p.dh = take(T * 512 * 2);
  p.dhn = take(T * OSD_HIDP * 2);
  p.dvg = take(T * 2 * OSD_HIDP * 2);
  p.dz2 = take(T * 512 * 2);
  p.dy = take(T * 1024 * 2);
  p.dys = take(T * 1024 * 2);  // w-scaled dO of the single-pass attention backward
  p.dqkv = take(T * 3072 * 2);
  p.dq_acc = take(T * 1024 * 4);  // fp32 dQ accumulator of the single-pass attention backward
  p.dz = take(T * 512 * 2);
  p.dsum = take(attn_bwd_fused_stats_floats(B, L, 16) * 4);  // >= T*16*4 (the two-pass kernels use the first B*16*L)
  p.da_tok = take(Ta * 128 * 4);
  p.a_pre = take(Ta * 128 * 4);
  p.da_pre = take(Ta * 128 * 2);
  p.audio_tm = take(Ta * 128 * 2);
  p.dcond = take(CondPack::floats(B, depth) * 4);
  p.dfsum = take((size_t)B * 64 * 4);
  p.dpre_s = take((size_t)B * 512 * 4 * 2);
  p.gWvg = take((size_t)2 * OSD_HIDP * 512 * 4);
  p.gWpo = take((size_t)512 * OSD_HIDP * 4);
  p.gbvg = take((size_t)2 * OSD_HIDP * 4);
  p.total = o;
  return p;
}

// OSD_ATTN_BWD=2pass selects the two-kernel attention backward (attn_bwd.cu) for A/B measurements
static bool attn_bwd_two_pass() {
  static const bool v = [] {
    const char* e = getenv("OSD_ATTN_BWD");
    return e != nullptr && strcmp(e, "2pass") == 0;
  }();
  return v;
}

// dX[M,N] = dY[M,K] * W[K,N]   (W stored [K rows][N] row-major = MN-major B operand)
static int gemm_dgrad(const void* dY, int64_t ldy, const void* W, int64_t ldw, void* dX, int64_t ldx, int x_fp32,
                      int atomic, int M, int N, int K, cudaStream_t s) {
  GemmArgs g;
  g.A = dY; g.a_major = MAJOR_K; g.lda = ldy; g.B = W; g.b_major = MAJOR_MN; g.ldb = ldw;
  g.M = M; g.N = N; g.K = K; g.elem = ELEM_BF16; g.epi = atomic ? EPI_ATOMIC : EPI_STORE; g.C = dX; g.ldc = ldx;
  g.c_fp32 = x_fp32; g.split_k = 1;
  return launch_gemm(g, s);
}
// dW[M,N] += dY[T,M]^T * X[T,N]   (both operands MN-major, contraction over tokens, split-K + fp32 atomics)
static int gemm_wgrad(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t ldw, int M, int N,
                      int T, cudaStream_t s) {
  GemmArgs g;
  g.A = dY; g.a_major = MAJOR_MN; g.lda = ldy; g.B = X; g.b_major = MAJOR_MN; g.ldb = ldx;
  g.M = M; g.N = N; g.K = T; g.elem = ELEM_BF16; g.epi = EPI_ATOMIC; g.C = dW; g.ldc = ldw; g.c_fp32 = 1;
  g.split_k = gemm_split_for(M, N, T);
  return launch_gemm(g, s);
}

static int pred_backward(const FwdCtx& c, const float* audio, const float* style, const float* xt, const float* du,
                         const float* dv, float* const* G, uint8_t* bw, cudaStream_t s) {
  const int B = c.B, L = c.L, T = B * L;
  OSD_CHECK(c.mode == OSD_BF16, "pred_backward: bf16 mode only");
  OSD_CHECK(c.a_batch == B, "pred_backward: broadcast audio (a_batch=1) is an inference-only shape");
  OSD_CHECK(c.save, "pred_backward: forward must have been run with save=1");
  const ActPlan& pl = c.plan;
  const int depth = c.depth, ts = tail_shift(depth);
  const BwdPlan bp = make_bwd_plan(B, L, c.a_batch, depth);
  auto LB = [&](int l) { return c.ws + (size_t)l * pl.layer_stride; };
  float* dx = reinterpret_cast<float*>(bw + bp.dx);
  void* dh = bw + bp.dh;
  float* dcond = reinterpret_cast<float*>(bw + bp.dcond);
  CondPack dc{dcond, B, depth};
  float* da_tok = reinterpret_cast<float*>(bw + bp.da_tok);
  float* gWvg = reinterpret_cast<float*>(bw + bp.gWvg);
  float* gWpo = reinterpret_cast<float*>(bw + bp.gWpo);
  float* gbvg = reinterpret_cast<float*>(bw + bp.gbvg);
  OSD_CUDA(cudaMemsetAsync(dcond, 0, CondPack::floats(B, depth) * 4, s));
  OSD_CUDA(cudaMemsetAsync(da_tok, 0, (size_t)T * 128 * 4, s));

  // ---- heads
  OSD_TRY(launch_final_bwd(reinterpret_cast<float*>(c.ws + pl.x_final), dv, c.P[P_OUT_W + ts], dx, G[P_OUT_W + ts],
                           G[P_OUT_B + ts], B, L, s));
  {
    float* dfsum = reinterpret_cast<float*>(bw + bp.dfsum);
    OSD_TRY(launch_u_final_bwd(du, reinterpret_cast<float*>(c.ws + pl.fsum), c.cond.umod(), c.P[P_UOUT_W + ts],
                               c.P[P_UOUT_B + ts], sqrtf(2.0f * OSD_E), L, dfsum, const_cast<float*>(dc.umod()),
                               G[P_UOUT_W + ts], G[P_UOUT_B + ts], B, s));
    const float* uw[8] = {c.P[P_UH0_W + ts], c.P[P_UH0_B + ts], c.P[P_UH1_W + ts], c.P[P_UH1_B + ts],
                          c.P[P_UH3_W + ts], c.P[P_UH3_B + ts], c.P[P_UH4_W + ts], c.P[P_UH4_B + ts]};
    float* ug[8] = {G[P_UH0_W + ts], G[P_UH0_B + ts], G[P_UH1_W + ts], G[P_UH1_B + ts],
                    G[P_UH3_W + ts], G[P_UH3_B + ts], G[P_UH4_W + ts], G[P_UH4_B + ts]};
    OSD_TRY(launch_u_head_bwd(xt, uw, dfsum, ug, B, L, s));
  }

  for (int l = depth - 1; l >= 0; --l) {
    uint8_t* lb = LB(l);
    const float* x0 = reinterpret_cast<float*>(lb + pl.x0);
    const float* x1 = reinterpret_cast<float*>(lb + pl.x1);
    float* dm1 = const_cast<float*>(dc.mod1(l));
    float* dm2 = const_cast<float*>(dc.mod2(l));
    // ---------------- ffn sub-block
    OSD_TRY(launch_postnorm_gate_bwd(dx, reinterpret_cast<float*>(lb + pl.f), c.cond.mod2(l), dh, dm2,
                                     G[lp(l, L_PO_B)], B, L, s));
    OSD_TRY(gemm_dgrad(dh, 512, c.W.po(l), OSD_HIDP, bw + bp.dhn, OSD_HIDP, 0, 0, T, OSD_HIDP, 512, s));
    OSD_CUDA(cudaMemsetAsync(gWpo, 0, (size_t)512 * OSD_HIDP * 4, s));
    OSD_TRY(gemm_wgrad(dh, 512, lb + pl.hn, OSD_HIDP, gWpo, OSD_HIDP, 512, OSD_HIDP, T, s));
    OSD_TRY(launch_unpack_grad(gWpo, G[lp(l, L_PO_W)], 512, OSD_HIDP, 512, OSD_HID, 0, 0, s));
    OSD_CUDA(cudaMemsetAsync(gbvg, 0, (size_t)2 * OSD_HIDP * 4, s));
    OSD_TRY(launch_swiglu_norm_bwd(lb + pl.vg, bw + bp.dhn, reinterpret_cast<float*>(lb + pl.rinv2), bw + bp.dvg, gbvg,
                                   T, s));
    OSD_TRY(launch_unpack_grad(gbvg, G[lp(l, L_VG_B)], 2 * OSD_HIDP, 1, 2 * OSD_HID, 1, OSD_HID, OSD_HIDP, s));
    OSD_TRY(gemm_dgrad(bw + bp.dvg, 2 * OSD_HIDP, c.W.vg(l), 512, bw + bp.dz2, 512, 0, 0, T, 512, 2 * OSD_HIDP, s));
    OSD_CUDA(cudaMemsetAsync(gWvg, 0, (size_t)2 * OSD_HIDP * 512 * 4, s));
    OSD_TRY(gemm_wgrad(bw + bp.dvg, 2 * OSD_HIDP, lb + pl.z2, 512, gWvg, 512, 2 * OSD_HIDP, 512, T, s));
    OSD_TRY(launch_unpack_grad(gWvg, G[lp(l, L_VG_W)], 2 * OSD_HIDP, 512, 2 * OSD_HID, 512, OSD_HID, OSD_HIDP, s));
    OSD_TRY(launch_dwconv_prenorm_bwd(bw + bp.dz2, lb + pl.hmod, x1, c.cond.mod2(l), c.P[lp(l, L_DW_W)], dx, dm2,
                                      G[lp(l, L_DW_W)], G[lp(l, L_DW_B)], B, L, s));
    // ---------------- attention sub-block
    OSD_TRY(launch_postnorm_gate_bwd(dx, reinterpret_cast<float*>(lb + pl.o), c.cond.mod1(l), dh, dm1,
                                     G[lp(l, L_OUT_B)], B, L, s));
    OSD_TRY(gemm_dgrad(dh, 512, c.W.out(l), 1024, bw + bp.dy, 1024, 0, 0, T, 1024, 512, s));
    OSD_TRY(gemm_wgrad(dh, 512, lb + pl.y, 1024, G[lp(l, L_OUT_W)], 1024, 512, 1024, T, s));
    const float* dq_acc = nullptr;  // single-pass path: dq stays in its fp32 accumulator, consumed by qknorm_rope_bwd
    const int* dq_flag = nullptr;
    if (attn_bwd_two_pass()) {
      OSD_TRY(launch_attn_bwd(lb + pl.qkv, lb + pl.y, bw + bp.dy, reinterpret_cast<float*>(lb + pl.lse),
                              reinterpret_cast<float*>(bw + bp.dsum), bw + bp.dqkv, B, L, 16, s));
    } else {
      OSD_TRY(launch_attn_bwd_fused(lb + pl.qkv, lb + pl.y, bw + bp.dy, reinterpret_cast<float*>(lb + pl.lse),
                                    reinterpret_cast<float*>(bw + bp.dsum), reinterpret_cast<float*>(bw + bp.dq_acc),
                                    bw + bp.dys, bw + bp.dqkv, B, L, 16, /*convert_dq=*/0, s));
      dq_acc = reinterpret_cast<const float*>(bw + bp.dq_acc);
      dq_flag = attn_bwd_fused_flag(reinterpret_cast<const float*>(bw + bp.dsum), B, L, 16);
    }
    OSD_TRY(launch_qknorm_rope_bwd(bw + bp.dqkv, lb + pl.qkv_raw, c.rope, c.P[lp(l, L_QN_W)], c.P[lp(l, L_KN_W)],
                                   G[lp(l, L_QN_W)], G[lp(l, L_KN_W)], G[lp(l, L_QKV_B)], dq_acc, dq_flag,
                                   attn_bwd_fused_dq_scale(), B, L, s));
    OSD_TRY(gemm_dgrad(bw + bp.dqkv, 3072, c.W.qkv(l), 512, bw + bp.dz, 512, 0, 0, T, 512, 3072, s));
    OSD_TRY(gemm_wgrad(bw + bp.dqkv, 3072, lb + pl.z, 512, G[lp(l, L_QKV_W)], 512, 3072, 512, T, s));
    OSD_TRY(launch_prenorm_mod_bwd(bw + bp.dz, x0, c.cond.mod1(l), dx, dm1, G[lp(l, L_CL_B)], B, L, s));
    OSD_TRY(gemm_wgrad(bw + bp.dz, 512, c.a_tok, 128, G[lp(l, L_CL_W)], 128, 512, 128, T, s));
    OSD_TRY(gemm_dgrad(bw + bp.dz, 512, c.W.cl(l), 128, da_tok, 128, 1, 1, T, 128, 512, s));
  }
  OSD_TRY(launch_proj_in_bwd(dx, xt, G[P_IN_W], G[P_IN_B], B, L, s));

  // ---- audio branch: a = silu(pre), pre = audio_tm Wa^T + ba
  {
    OSD_TRY(launch_cf_to_tm(audio, bw + bp.audio_tm, 1, B, 128, L, s));
    GemmArgs g;
    g.A = bw + bp.audio_tm; g.B = c.W.wa(); g.lda = 128; g.ldb = 128; g.M = T; g.N = 128; g.K = 128;
    g.elem = ELEM_BF16; g.epi = EPI_STORE; g.C = bw + bp.a_pre; g.ldc = 128; g.c_fp32 = 1; g.bias = c.P[P_AUDIO_B];
    OSD_TRY(launch_gemm(g, s));
    OSD_TRY(launch_silu_bwd(da_tok, reinterpret_cast<float*>(bw + bp.a_pre), bw + bp.da_pre, G[P_AUDIO_B], T, s));
    OSD_TRY(gemm_wgrad(bw + bp.da_pre, 128, bw + bp.audio_tm, 128, G[P_AUDIO_W], 128, 128, 128, T, s));
  }
  // ---- conditioning vectors (fp32, tiny)
  float* dcg = dcond;  // [B,512]
  float* scratch = reinterpret_cast<float*>(bw + bp.dpre_s);
  for (int l = 0; l < depth; ++l) {
    OSD_TRY(launch_linear_small_bwd(dc.mod1(l), nullptr, c.cond.cg(), c.P[lp(l, L_SSG1_W)], G[lp(l, L_SSG1_W)],
                                    G[lp(l, L_SSG1_B)], dcg, nullptr, B, 1536, 512, 0, s));
    OSD_TRY(launch_linear_small_bwd(dc.mod2(l), nullptr, c.cond.cg(), c.P[lp(l, L_SSG2_W)], G[lp(l, L_SSG2_W)],
                                    G[lp(l, L_SSG2_B)], dcg, nullptr, B, 1536, 512, 0, s));
  }
  OSD_TRY(launch_linear_small_bwd(dc.umod(), nullptr, c.cond.cg(), c.P[P_UMOD_W + ts], G[P_UMOD_W + ts], G[P_UMOD_B + ts],
                                  dcg, nullptr, B, 128, 512, 0, s));
  // cg = silu(pre_s): recompute the pre-activation, then the layer's own gradients
  float* pre_s = scratch;
  float* dpre_s = scratch + (size_t)B * 512;
  OSD_TRY(launch_linear_small(style, c.P[P_STYLE_W], c.P[P_STYLE_B], pre_s, B, 512, 32, 0, s));
  OSD_TRY(launch_linear_small_bwd(dcg, pre_s, style, c.P[P_STYLE_W], G[P_STYLE_W], G[P_STYLE_B], nullptr, dpre_s, B, 512,
                                  32, 1, s));
  return 0;
}

}  // namespace osd

// ================================================================================================
using namespace osd;

extern "C" {

// every `mode` argument: precision in bits 0-7, backbone depth in bits 8-15 (0 = OSD_DEPTH), see OSD_MODE()
#define OSD_SPLIT_MODE(fn)                                                                              \
  const int depth = depth_of(mode);                                                                     \
  mode = prec_of(mode);                                                                                 \
  OSD_CHECK(mode == OSD_BF16 || mode == OSD_F32X3, fn ": bad precision %d", mode);                      \
  OSD_CHECK(depth >= 1 && depth <= OSD_MAX_DEPTH, fn ": depth %d not in 1..%d", depth, OSD_MAX_DEPTH)

static FwdCtx make_ctx(const float* const* params, const void* packed, int mode, int depth, const void* a_tok,
                       const float* cond, const float* rope, int B, int L, int a_batch, void* workspace, int save) {
  FwdCtx c;
  c.P = params;
  c.W = PackedW{static_cast<const uint8_t*>(packed), packed_layout(mode, depth)};
  c.mode = mode; c.depth = depth; c.B = B; c.L = L; c.a_batch = a_batch;
  c.cond = CondPack{cond, B, depth};
  c.a_tok = a_tok;
  c.rope = rope;
  c.ws = static_cast<uint8_t*>(workspace);
  c.plan = make_plan(B, L, a_batch, mode, save, depth);
  c.save = save;
  c.cl_hoisted = nullptr;
  return c;
}

int osd_num_params(int mode) { return num_params(depth_of(mode)); }
size_t osd_packed_bytes(int mode) { return packed_layout(prec_of(mode), depth_of(mode)).total; }
size_t osd_cond_floats(int B) { return CondPack::floats(B, OSD_DEPTH); }
size_t osd_cond_floats_mode(int B, int mode) { return CondPack::floats(B, depth_of(mode)); }
size_t osd_workspace_bytes(int B, int L, int a_batch, int mode, int save) {
  return make_plan(B, L, a_batch, prec_of(mode), save, depth_of(mode)).total;
}
static size_t sample_extra_bytes(int B, int L, int a_batch, int depth) {
  // hoisted proj_cl outputs [depth][Ta,512] bf16 / fp32 + v [B,6,L] + u [B] + eta [2]
  return depth * al((size_t)a_batch * L * 512 * 4) + al((size_t)B * 6 * L * 4) + al((size_t)B * 4) + 1024;
}
size_t osd_sample_extra_bytes(int B, int L, int a_batch) { return sample_extra_bytes(B, L, a_batch, OSD_DEPTH); }
size_t osd_sample_extra_bytes_mode(int B, int L, int a_batch, int mode) {
  return sample_extra_bytes(B, L, a_batch, depth_of(mode));
}

int osd_pack_weights(const float* const* params, void* packed, int mode, void* stream) {
  OSD_CHECK(params && packed, "osd_pack_weights: null argument");
  OSD_SPLIT_MODE("osd_pack_weights");
  return pack_weights(params, static_cast<uint8_t*>(packed), mode, depth, static_cast<cudaStream_t>(stream));
}

int osd_precompute_conditioning(const float* const* params, const void* packed, int mode, const float* audio,
                                int a_batch, const float* style, int B, int L, void* scratch, void* a_tok,
                                float* cond, void* stream) {
  OSD_CHECK(params && packed && audio && style && scratch && a_tok && cond,
            "osd_precompute_conditioning: null argument");
  OSD_CHECK(a_batch == B || a_batch == 1, "osd_precompute_conditioning: audio batch %d must be 1 or B=%d", a_batch, B);
  OSD_SPLIT_MODE("osd_precompute_conditioning");
  PackedW W{static_cast<const uint8_t*>(packed), packed_layout(mode, depth)};
  return precompute_conditioning(params, W, mode, depth, audio, a_batch, style, B, L, a_tok, cond, scratch,
                                 static_cast<cudaStream_t>(stream));
}

// The (a, cg) pair of DiffusionModel._pred (model.py:86-103) given as the reference passes it -- a [a_batch,128,L] and
// cg [B,512], fp32 channels-first -- turned into what osd_pred_forward consumes: a_tok (token-major operand, (hi | lo)
// pairs in the fp32-grade mode) and the cond pack (cg | modulation vectors of every layer | u_mod).
int osd_conditioning_from(const float* const* params, int mode, const float* a, int a_batch, const float* cg, int B, int L,
                          void* a_tok, float* cond, void* stream) {
  OSD_CHECK(params && a && cg && a_tok && cond, "osd_conditioning_from: null argument");
  OSD_CHECK(a_batch == B || a_batch == 1, "osd_conditioning_from: audio batch %d must be 1 or B=%d", a_batch, B);
  OSD_SPLIT_MODE("osd_conditioning_from");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == OSD_F32X3)
    OSD_TRY(launch_cf_to_tm_split(a, a_tok, a_batch, 128, L, s));
  else
    OSD_TRY(launch_cf_to_tm(a, a_tok, 1, a_batch, 128, L, s));
  if (cg != cond) OSD_CUDA(cudaMemcpyAsync(cond, cg, (size_t)B * 512 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return modulation_from_cg(params, depth, cond, B, s);
}

int osd_tokens_to_channels(const void* in, int in_fp32, float* out, int B, int C, int L, void* stream) {
  return launch_tm_to_cf(in, in_fp32, out, B, C, L, static_cast<cudaStream_t>(stream));
}
int osd_channels_to_tokens(const float* in, void* out, int out_fp32, int B, int C, int L, void* stream) {
  return launch_cf_to_tm(in, out, !out_fp32, B, C, L, static_cast<cudaStream_t>(stream));
}

int osd_pred_forward(const float* const* params, const void* packed, int mode, const void* a_tok, const float* cond,
                     const float* rope, const float* xt, float* u, float* v, int B, int L, int a_batch,
                     void* workspace, int save, void* stream) {
  OSD_CHECK(params && packed && a_tok && cond && rope && xt && u && v && workspace, "osd_pred_forward: null argument");
  OSD_SPLIT_MODE("osd_pred_forward");
  const FwdCtx c = make_ctx(params, packed, mode, depth, a_tok, cond, rope, B, L, a_batch, workspace, save);
  return pred_forward(c, xt, u, v, static_cast<cudaStream_t>(stream));
}

// DiffusionModel.sample (model.py:117-138): x is the initial noise on entry and the sample on exit.
// The step-invariant proj_cl(a) of all 8 layers is computed once; the u0 probe stays on the device.
int osd_sample(const float* const* params, const void* packed, int mode, const void* a_tok, const float* cond,
               const float* rope, float* x, int num_steps, float c0, int B, int L, int a_batch, void* workspace,
               void* extra, float* eta_u0_out, void* stream) {
  OSD_CHECK(params && packed && a_tok && cond && rope && x && workspace && extra, "osd_sample: null argument");
  OSD_CHECK(num_steps >= 1, "osd_sample: num_steps must be >= 1");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OSD_SPLIT_MODE("osd_sample");
  FwdCtx c = make_ctx(params, packed, mode, depth, a_tok, cond, rope, B, L, a_batch, workspace, 0);
  uint8_t* ex = static_cast<uint8_t*>(extra);
  const int Ta = a_batch * L;
  const size_t cl_bytes = al((size_t)Ta * 512 * (mode == OSD_F32X3 ? 4 : 2));
  for (int l = 0; l < depth; ++l) OSD_TRY(proj_cl(params, c.W, mode, l, a_tok, Ta, ex + l * cl_bytes, s));
  c.cl_hoisted = ex;
  float* v = reinterpret_cast<float*>(ex + depth * cl_bytes);
  float* u = reinterpret_cast<float*>(ex + depth * cl_bytes + al((size_t)B * 6 * L * 4));
  float* eta = reinterpret_cast<float*>(ex + depth * cl_bytes + al((size_t)B * 6 * L * 4) + al((size_t)B * 4));
  OSD_TRY(pred_forward(c, x, u, v, s));
  OSD_TRY(launch_sample_eta(u, B, sqrtf(c0), num_steps, eta, s));
  for (int i = 0; i < num_steps; ++i) {
    OSD_TRY(pred_forward(c, x, u, v, s));
    OSD_TRY(launch_sample_update(x, v, u, eta, B, L, s));
  }
  if (eta_u0_out != nullptr)
    OSD_CUDA(cudaMemcpyAsync(eta_u0_out, eta, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}


size_t osd_backward_workspace_bytes(int B, int L, int a_batch) { return make_bwd_plan(B, L, a_batch, OSD_MAX_DEPTH).total; }

// Gradient of osd_pred_forward(save=1) composed with osd_precompute_conditioning: given du [B], dv [B,6,L]
// ACCUMULATES d(loss)/d(parameter) into grads[164] (fp32, parameter shapes, caller zero-initialised or
// carrying earlier accumulation).
int osd_pred_backward(const float* const* params, const void* packed, int mode, const void* a_tok, const float* cond,
                      const float* rope, const float* audio, const float* style, const float* xt, const float* du,
                      const float* dv, float* const* grads, int B, int L, int a_batch, void* workspace,
                      void* bwd_workspace, void* stream) {
  OSD_CHECK(params && packed && a_tok && cond && rope && audio && style && xt && du && dv && grads && workspace &&
                bwd_workspace,
            "osd_pred_backward: null argument");
  OSD_SPLIT_MODE("osd_pred_backward");
  const FwdCtx c = make_ctx(params, packed, mode, depth, a_tok, cond, rope, B, L, a_batch, workspace, 1);
  return pred_backward(c, audio, style, xt, du, dv, grads, static_cast<uint8_t*>(bwd_workspace),
                       static_cast<cudaStream_t>(stream));
}

int osd_attn_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* dsum, void* dqkv, int B, int L,
                 int H, void* stream) {
  return launch_attn_bwd(qkv, y, dy, lse, dsum, dqkv, B, L, H, static_cast<cudaStream_t>(stream));
}

// debugging aid (tools/trace_attn_bwd.py): event timeline of one CTA of the single-pass kernel; buf = device
// buffer of 3 x 1024 u64 records, or null to switch tracing off
void osd_debug_attn_bwd_trace(unsigned long long* buf, int cta) { attn_bwd_fused_set_trace(buf, cta); }

size_t osd_attn_bwd_fused_stats_floats(int B, int L, int H) { return attn_bwd_fused_stats_floats(B, L, H); }

int osd_attn_bwd_fused(const void* qkv, const void* y, const void* dy, const float* lse, float* stats, float* dq_acc,
                       void* dy_scaled, void* dqkv, int B, int L, int H, void* stream) {
  return launch_attn_bwd_fused(qkv, y, dy, lse, stats, dq_acc, dy_scaled, dqkv, B, L, H, /*convert_dq=*/1,
                               static_cast<cudaStream_t>(stream));
}

}  // extern "C"
