// Denoiser orchestration: parameter indexing, packed operand weights, activation plans.
#pragma once
#include "../../include/osd_b200.h"
#include "common.h"

namespace osd {

// ---- parameter indices in reference state-dict order (164 tensors) ----
enum {
  P_AUDIO_W = 0, P_AUDIO_B, P_STYLE_W, P_STYLE_B, P_IN_W, P_IN_B,
  P_LAYER0 = 6, P_LAYER_STRIDE = 18,
  L_SSG1_W = 0, L_SSG1_B, L_CL_W, L_CL_B, L_QKV_W, L_QKV_B, L_OUT_W, L_OUT_B, L_QN_W, L_KN_W,
  L_SSG2_W, L_SSG2_B, L_DW_W, L_DW_B, L_VG_W, L_VG_B, L_PO_W, L_PO_B,
  P_OUT_W = 150, P_OUT_B, P_UH0_W, P_UH0_B, P_UH1_W, P_UH1_B, P_UH3_W, P_UH3_B, P_UH4_W, P_UH4_B,
  P_UMOD_W, P_UMOD_B, P_UOUT_W, P_UOUT_B,
  P_COUNT = 164
};
inline int lp(int layer, int which) { return P_LAYER0 + layer * P_LAYER_STRIDE + which; }
// `mode` arguments of the C ABI: bits 0-7 precision (OSD_BF16 / OSD_F32X3), bits 8-15 backbone depth (0 = OSD_DEPTH);
// the tail tensors (proj_out ... u_out) follow the last layer: P_OUT_W + tail_shift(depth)
inline int prec_of(int mode) { return mode & 0xff; }
inline int depth_of(int mode) { const int d = (mode >> 8) & 0xff; return d ? d : OSD_DEPTH; }
inline int tail_shift(int depth) { return (depth - OSD_DEPTH) * P_LAYER_STRIDE; }
inline int num_params(int depth) { return P_COUNT + tail_shift(depth); }

// ---- packed operand weights (bf16 or fp32-for-tf32), one buffer ----
struct PackedLayout {
  size_t esz;  // operand element size (bf16)
  int kmul;    // 2 in the split-precision mode: every weight row is (hi | lo)
  size_t wa;   // [128,128]
  size_t layer0, layer_stride;
  size_t l_cl, l_qkv, l_out, l_vg, l_po;  // offsets inside a layer block (bytes)
  size_t bvg;                             // fp32 [depth][2816]
  size_t bounds;                          // fp32 [depth] attention score bounds (log2 units) per layer
  size_t total;
};
PackedLayout packed_layout(int mode, int depth);

struct PackedW {
  const uint8_t* base;
  PackedLayout lay;
  const void* wa() const { return base + lay.wa; }
  const void* cl(int l) const { return base + lay.layer0 + l * lay.layer_stride + lay.l_cl; }
  const void* qkv(int l) const { return base + lay.layer0 + l * lay.layer_stride + lay.l_qkv; }
  const void* out(int l) const { return base + lay.layer0 + l * lay.layer_stride + lay.l_out; }
  const void* vg(int l) const { return base + lay.layer0 + l * lay.layer_stride + lay.l_vg; }
  const void* po(int l) const { return base + lay.layer0 + l * lay.layer_stride + lay.l_po; }
  const float* bound(int l) const { return reinterpret_cast<const float*>(base + lay.bounds) + l; }
  const float* bvg(int l) const { return reinterpret_cast<const float*>(base + lay.bvg) + (size_t)l * 2 * OSD_HIDP; }
};

// ---- conditioning pack (fp32): cg [B,512] | mod1 [depth][B,1536] | mod2 [depth][B,1536] | umod [B,128] ----
struct CondPack {
  const float* base;
  int B;
  int depth;
  const float* cg() const { return base; }
  const float* mod1(int l) const { return base + (size_t)B * 512 + (size_t)l * B * 1536; }
  const float* mod2(int l) const { return base + (size_t)B * 512 + (size_t)(depth + l) * B * 1536; }
  const float* umod() const { return base + (size_t)B * 512 + (size_t)2 * depth * B * 1536; }
  static size_t floats(int B, int depth) { return (size_t)B * (512 + (size_t)2 * depth * 1536 + 128); }
};

// ---- activation plan: per-layer buffers; layer stride 0 (inference, buffers reused) or >0 (training saves) ----
struct ActPlan {
  // token-major; "op" = operand dtype (bf16 / fp32)
  size_t x0, cl, z, qkv_raw, qkv, y, lse, o, x1, hmod, z2, vg, hn, rinv2, f;  // offsets inside one layer block
  size_t layer_bytes;   // bytes of one layer block
  size_t layer_stride;  // 0 when buffers are shared by all layers
  size_t x_final;       // fp32 [T,512] output of the last layer (after the layer blocks)
  size_t fsum;          // fp32 [B,64]
  size_t fpart;         // fp32 [B, u_head blocks, 64]: per-block partial sums of the u head (deterministic mean)
  size_t uh1, uh2;      // fp32 [T,64] u-head saves (training only)
  size_t total;
};
ActPlan make_plan(int B, int L, int a_batch, int mode, int save, int depth);

}  // namespace osd
