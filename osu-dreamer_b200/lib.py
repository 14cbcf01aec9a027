"""ctypes binding of libosd_b200.so (C ABI: include/osd_b200.h).

torch is used only as the owner of device memory and streams: tensors are passed as raw device
pointers, work is enqueued on torch's current CUDA stream.  There is no CPU fallback: if the shared
library is missing or a tensor is not on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OSD_LIB_PATH', os.path.join(_HERE, 'libosd_b200.so'))  # override: A/B kernel experiments
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'osd_b200.h')

_lib = None

BF16, TF32 = 0, 1   # operand element kinds of osd_gemm
MODE_BF16, MODE_F32X3 = 0, 1  # model precision modes (include/osd_b200.h)
DEFAULT_DEPTH, MAX_DEPTH = 8, 32


def mode_of(precision: int, depth: int = DEFAULT_DEPTH) -> int:
    """OSD_MODE(precision, depth): every `mode` argument of the C ABI carries the backbone depth in bits 8-15."""
    return precision | (depth << 8)


def depth_of(mode: int) -> int:
    return ((mode >> 8) & 0xff) or DEFAULT_DEPTH
MAJOR_K, MAJOR_MN = 0, 1
EPI_STORE, EPI_SILU, EPI_ATOMIC = 0, 1, 2


class OsdError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises OsdError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OsdError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       f'or `make -C osu-dreamer_b200/csrc` (there is no CPU fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    lib.osd_last_error.restype = c_char_p
    lib.osd_abi_version.restype = c_int
    _lib = lib
    return lib


def _check(status: int):
    if status != 0:
        msg = load().osd_last_error()
        raise OsdError(f'libosd_b200 status {status}: {msg.decode() if msg else "?"}')


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise OsdError('libosd_b200 has no CPU path: tensor must live on a CUDA device')
    if t.device.index != torch.cuda.current_device():
        raise OsdError(f'tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}: the work would be '
                       f'enqueued on the wrong device\'s stream (wrap the call in `with torch.cuda.device(t.device)`)')
    return c_void_p(t.data_ptr())


def stream():
    """torch's current stream on the CURRENT device: tensors handed to the library must live there (see `ptr`)."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _elem_of(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return TF32
    raise OsdError(f'unsupported operand dtype {t.dtype}')


def gemm(A, B, C, bias=None, a_major=MAJOR_K, b_major=MAJOR_K, epi=EPI_STORE, split_k=1):
    """C[M,N] (+)= A * B^T.  A: [M,K] (MAJOR_K) or [K,M] (MAJOR_MN); B: [N,K] or [K,N]; 2-D, inner-contiguous."""
    assert A.dim() == 2 and B.dim() == 2 and C.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
    M, K = (A.shape if a_major == MAJOR_K else (A.shape[1], A.shape[0]))
    N, Kb = (B.shape if b_major == MAJOR_K else (B.shape[1], B.shape[0]))
    assert K == Kb and C.shape == (M, N) and A.dtype == B.dtype
    _check(load().osd_gemm(ptr(A), c_int(a_major), c_int64(A.stride(0)), ptr(B), c_int(b_major), c_int64(B.stride(0)),
                           ptr(C), c_int64(C.stride(0)), c_int(1 if C.dtype == torch.float32 else 0), ptr(bias),
                           c_int(M), c_int(N), c_int(K), c_int(_elem_of(A)), c_int(epi), c_int(split_k), stream()))
    return C


def gemm_split_k(M: int, N: int, K: int) -> int:
    """the split-K factor the library's weight-gradient GEMMs use for this shape (whole waves of the persistent grid)"""
    return int(load().osd_gemm_split_k(c_int(M), c_int(N), c_int(K)))


def rope_table(L: int, device) -> torch.Tensor:
    """fp32 cos|sin table [L, 2, 32] followed by its 32-row-transposed copy (flat, osd_rope_table_floats(L) floats);
    inv_freq formed exactly as osu_dreamer/common/attn.py:16-18."""
    inv_freq = (10000 ** (torch.arange(0, 64, 2).float() / -64)).contiguous()
    arr = (c_float * 32)(*inv_freq.tolist())
    out = torch.empty(_sz('osd_rope_table_floats', L), dtype=torch.float32, device=device)
    _check(load().osd_rope_table(arr, c_int(L), ptr(out), stream()))
    return out


def qkv_proj(x, w, bias, qnorm_w, knorm_w, rope, L, raw_out=None):
    T = x.shape[0]
    out = torch.empty(T, 3072, dtype=torch.bfloat16, device=x.device)
    _check(load().osd_qkv_proj(ptr(x), ptr(w), ptr(bias), ptr(qnorm_w), ptr(knorm_w), ptr(rope), ptr(out),
                               ptr(raw_out), c_int(T), c_int(L), c_int(_elem_of(x)), stream()))
    return out


# ---------------------------------------------------------------------------------------------
# model-level entry points
# ---------------------------------------------------------------------------------------------
NUM_PARAMS = 164


def _sz(fn, *args):
    lib = load()
    f = getattr(lib, fn)
    f.restype = c_size_t
    return int(f(*[c_int(a) for a in args]))


def packed_bytes(mode):
    return _sz('osd_packed_bytes', mode)


def cond_floats(B, mode=0):
    return _sz('osd_cond_floats_mode', B, mode)


def num_params(mode=0):
    return 20 + 18 * depth_of(mode)


def workspace_bytes(B, L, a_batch, mode, save):
    return _sz('osd_workspace_bytes', B, L, a_batch, mode, save)


def sample_extra_bytes(B, L, a_batch, mode=0):
    return _sz('osd_sample_extra_bytes_mode', B, L, a_batch, mode)


def param_array(tensors):
    """HOST array of device pointers, reference state-dict order (20 + 18 * depth tensors)."""
    if (len(tensors) - 20) % 18 or not 1 <= (len(tensors) - 20) // 18 <= MAX_DEPTH:
        raise OsdError(f'{len(tensors)} parameter tensors: expected 20 + 18 * depth')
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise OsdError('parameters must be contiguous fp32 CUDA tensors (no CPU path)')
        if t.data_ptr() % 16:
            raise OsdError('parameter storage must be 16-byte aligned (vector loads)')
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def attn_fwd(qkv, B, L, H=16, want_lse=True, bound_log2=None, variant=7):
    T = B * L
    y = torch.empty(T, H * 64, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, L, dtype=torch.float32, device=qkv.device) if want_lse else None
    _check(load().osd_attn_fwd(ptr(qkv), ptr(y), ptr(lse), ptr(bound_log2), c_int(B), c_int(L), c_int(H),
                               c_int(variant), stream()))
    return y, lse


def pack_weights(parr, packed, mode):
    _check(load().osd_pack_weights(parr, ptr(packed), c_int(mode), stream()))


def precompute_conditioning(parr, packed, mode, audio, style, scratch, a_tok, cond):
    a_batch, _, L = audio.shape
    B = style.shape[0]
    _check(load().osd_precompute_conditioning(parr, ptr(packed), c_int(mode), ptr(audio), c_int(a_batch), ptr(style),
                                              c_int(B), c_int(L), ptr(scratch), ptr(a_tok), ptr(cond), stream()))


def conditioning_from(parr, mode, a, cg, L, a_tok, cond):
    """(a [#B,128,L], cg [B,512]) as DiffusionModel._pred receives them -> (a_tok, cond pack) for pred_forward."""
    _check(load().osd_conditioning_from(parr, c_int(mode), ptr(a), c_int(a.shape[0]), ptr(cg), c_int(cg.shape[0]),
                                        c_int(L), ptr(a_tok), ptr(cond), stream()))


def pred_forward(parr, packed, mode, a_tok, cond, rope, xt, u, v, a_batch, workspace, save):
    B, _, L = xt.shape
    _check(load().osd_pred_forward(parr, ptr(packed), c_int(mode), ptr(a_tok), ptr(cond), ptr(rope), ptr(xt), ptr(u),
                                   ptr(v), c_int(B), c_int(L), c_int(a_batch), ptr(workspace), c_int(save), stream()))


def sample(parr, packed, mode, a_tok, cond, rope, x, num_steps, c0, a_batch, workspace, extra, eta_u0):
    B, _, L = x.shape
    _check(load().osd_sample(parr, ptr(packed), c_int(mode), ptr(a_tok), ptr(cond), ptr(rope), ptr(x),
                             c_int(num_steps), c_float(c0), c_int(B), c_int(L), c_int(a_batch), ptr(workspace),
                             ptr(extra), ptr(eta_u0), stream()))


def tokens_to_channels(tok, B, C, L):
    out = torch.empty(B, C, L, dtype=torch.float32, device=tok.device)
    _check(load().osd_tokens_to_channels(ptr(tok), c_int(1 if tok.dtype == torch.float32 else 0), ptr(out), c_int(B),
                                         c_int(C), c_int(L), stream()))
    return out


def channels_to_tokens(x, out_dtype=torch.bfloat16):
    B, C, L = x.shape
    out = torch.empty(B * L, C, dtype=out_dtype, device=x.device)
    _check(load().osd_channels_to_tokens(ptr(x), ptr(out), c_int(1 if out_dtype == torch.float32 else 0), c_int(B),
                                         c_int(C), c_int(L), stream()))
    return out


def backward_workspace_bytes(B, L, a_batch):
    return _sz('osd_backward_workspace_bytes', B, L, a_batch)


def grad_array(tensors):
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()) or t.data_ptr() % 16:
            raise OsdError('gradient buffers must be contiguous, 16-byte aligned fp32 CUDA tensors')
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def pred_backward(parr, packed, mode, a_tok, cond, rope, audio, style, xt, du, dv, garr, a_batch, workspace, bwd_ws):
    B, _, L = xt.shape
    _check(load().osd_pred_backward(parr, ptr(packed), c_int(mode), ptr(a_tok), ptr(cond), ptr(rope), ptr(audio),
                                    ptr(style), ptr(xt), ptr(du), ptr(dv), garr, c_int(B), c_int(L), c_int(a_batch),
                                    ptr(workspace), ptr(bwd_ws), stream()))


def attn_bwd(qkv, y, dy, lse, B, L, H=16):
    dqkv = torch.empty_like(qkv)
    dsum = torch.empty(B, H, L, dtype=torch.float32, device=qkv.device)
    _check(load().osd_attn_bwd(ptr(qkv), ptr(y), ptr(dy), ptr(lse), ptr(dsum), ptr(dqkv), c_int(B), c_int(L), c_int(H),
                               stream()))
    return dqkv


def attn_bwd_fused(qkv, y, dy, lse, B, L, H=16):
    dqkv = torch.empty_like(qkv)
    stats = torch.empty(_sz('osd_attn_bwd_fused_stats_floats', B, L, H), dtype=torch.float32,
                        device=qkv.device)
    dq_acc = torch.empty(B * L, H * 64, dtype=torch.float32, device=qkv.device)
    dys = torch.empty_like(dy)
    _check(load().osd_attn_bwd_fused(ptr(qkv), ptr(y), ptr(dy), ptr(lse), ptr(stats), ptr(dq_acc), ptr(dys), ptr(dqkv), c_int(B),
                                     c_int(L), c_int(H), stream()))
    return dqkv


def adamw_ema_step(p, g, m, v, ema, step, lr, beta1, beta2, eps, wd, max_norm, grad_scale, ema_decay, ema_copy, acc,
                   scal):
    _check(load().osd_adamw_ema_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), c_size_t(p.numel()), c_int(step),
                                     c_float(lr), c_float(beta1), c_float(beta2), c_float(eps), c_float(wd),
                                     c_float(max_norm), c_float(grad_scale), c_float(ema_decay),
                                     c_int(1 if ema_copy else 0), ptr(acc), ptr(scal), stream()))


def launch_count() -> int:
    f = load().osd_launch_count
    f.restype = ctypes.c_ulonglong
    return int(f())


def loss_fwd_bwd(xt, x1, u, v, c0, osl_w, del_w):
    """-> (out4 = [loss, osl, del, u_mape], du [B], dv [B,6,L]) -- fused loss value + output gradients."""
    B, _, L = xt.shape
    dev = xt.device
    out4 = torch.empty(4, dtype=torch.float32, device=dev)
    du = torch.empty(B, dtype=torch.float32, device=dev)
    dv = torch.empty_like(v)
    scratch = torch.empty(4 * B, dtype=torch.float32, device=dev)
    _check(load().osd_loss_fwd_bwd(ptr(xt), ptr(x1), ptr(u), ptr(v), c_float(c0), c_float(osl_w), c_float(del_w),
                                   c_int(B), c_int(L), ptr(out4), ptr(du), ptr(dv), ptr(scratch), stream()))
    return out4, du, dv


# ------------------------------------------------------------------ style model inference (csrc/style.cu)
STYLE_NUM_PARAMS = 60


def style_param_array(tensors):
    """HOST array of device pointers, reference state-dict order (parameters and the two Fourier-feature buffers)."""
    if len(tensors) != STYLE_NUM_PARAMS:
        raise OsdError(f'style model: expected {STYLE_NUM_PARAMS} tensors, got {len(tensors)}')
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()) or t.data_ptr() % 16:
            raise OsdError('style model tensors must be contiguous, 16-byte aligned fp32 CUDA tensors (no CPU path)')
    return (c_void_p * STYLE_NUM_PARAMS)(*[t.data_ptr() for t in tensors])


def style_forward(parr, st, labels):
    B = st.shape[0]
    u = torch.empty(B, dtype=torch.float32, device=st.device)
    v = torch.empty(B, st.shape[1], dtype=torch.float32, device=st.device)
    scratch = torch.empty(_sz('osd_style_scratch_floats', B), dtype=torch.float32, device=st.device)
    _check(load().osd_style_forward(parr, ptr(st), ptr(labels), ptr(u), ptr(v), ptr(scratch), c_int(B), stream()))
    return u, v


def style_sample(parr, labels, s, num_steps):
    """in place on s [B, 32]; returns the {eta, u0} pair (device tensor)."""
    B = s.shape[0]
    scratch = torch.empty(_sz('osd_style_scratch_floats', B), dtype=torch.float32, device=s.device)
    eta_u0 = torch.empty(2, dtype=torch.float32, device=s.device)
    _check(load().osd_style_sample(parr, ptr(labels), ptr(s), c_int(num_steps), ptr(scratch), ptr(eta_u0), c_int(B), stream()))
    return eta_u0


# ------------------------------------------------------------------ style model training (csrc/style_train.cu)
def style_train_workspace(B, device):
    return torch.empty(_sz('osd_style_train_workspace_floats', B), dtype=torch.float32, device=device)


def style_train_forward(parr, st, labels, ws):
    """StyleModel.forward keeping its activations in ws -> (u [B], v [B,32])."""
    B = st.shape[0]
    u = torch.empty(B, dtype=torch.float32, device=st.device)
    v = torch.empty(B, st.shape[1], dtype=torch.float32, device=st.device)
    _check(load().osd_style_train_forward(parr, ptr(st), ptr(labels), ptr(u), ptr(v), ptr(ws), c_int(B), stream()))
    return u, v


def style_loss(st, s1, u, v, osl_w, del_w):
    """-> (out4 = [loss, osl, del, u_mape], du [B], dv [B,32]): StyleTrainer.forward's loss value + output gradients."""
    B = st.shape[0]
    out4 = torch.empty(4, dtype=torch.float32, device=st.device)
    du = torch.empty(B, dtype=torch.float32, device=st.device)
    dv = torch.empty_like(v)
    acc = torch.empty(4, dtype=torch.float32, device=st.device)
    _check(load().osd_style_loss(ptr(st), ptr(s1), ptr(u), ptr(v), c_float(osl_w), c_float(del_w), ptr(out4), ptr(du), ptr(dv),
                                 ptr(acc), c_int(B), stream()))
    return out4, du, dv


def style_grad_array(tensors):
    """HOST array of 60 device pointers: gradient buffers in state-dict order; None (the two Fourier-feature buffers) -> NULL."""
    if len(tensors) != STYLE_NUM_PARAMS:
        raise OsdError(f'style model: expected {STYLE_NUM_PARAMS} gradient slots, got {len(tensors)}')
    for t in tensors:
        if t is not None and (not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()) or t.data_ptr() % 16):
            raise OsdError('style gradient buffers must be contiguous, 16-byte aligned fp32 CUDA tensors')
    return (c_void_p * STYLE_NUM_PARAMS)(*[0 if t is None else t.data_ptr() for t in tensors])


def style_backward(parr, st, labels, du, dv, garr, ws):
    _check(load().osd_style_backward(parr, ptr(st), ptr(labels), ptr(du), ptr(dv), garr, ptr(ws), c_int(st.shape[0]), stream()))


# ------------------------------------------------------------------ latent model, inference half (csrc/latent.cu)
def _f32(t):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise OsdError('latent ops take contiguous fp32 CUDA tensors (no CPU path)')
    return t


def lat_block(x, w8, film=None):
    """one residual SwiGLU block (unet.py:50-54); x [B,128,L] -> new tensor; w8 = the block's 8 tensors; film [B,384] or None"""
    B, C, L = x.shape
    y = torch.empty_like(x)
    arr = (c_void_p * 8)(*[_f32(t).data_ptr() for t in w8])
    _check(load().osd_lat_block(ptr(_f32(x)), ptr(y), arr, ptr(_f32(film)) if film is not None else c_void_p(0), c_int(B), c_int(L), stream()))
    return y


def lat_tc_pack(w1, b1, w2):
    """(hi | lo) tf32-split copies of a block's two 1x1-conv weights + the padded bias, for lat_block_tc (written once per block)"""
    lib_ = load()
    lib_.osd_lat_tc_pack_bytes.restype = c_size_t
    packed = torch.empty(int(lib_.osd_lat_tc_pack_bytes()), dtype=torch.uint8, device=w1.device)
    _check(lib_.osd_lat_tc_pack(ptr(_f32(w1)), ptr(_f32(b1)), ptr(_f32(w2)), ptr(packed), stream()))
    return packed


def lat_tc_workspace(B, L, device):
    return torch.empty(_sz('osd_lat_tc_workspace_bytes', B, L), dtype=torch.uint8, device=device)


def lat_block_tc(x, w8, packed, film=None, ws=None):
    """lat_block with its two 1x1 convolutions as 3xTF32 tcgen05 GEMMs (~1e-6 of fp32); same arguments + the packed weights"""
    B, C, L = x.shape
    y = torch.empty_like(x)
    if ws is None:
        ws = lat_tc_workspace(B, L, x.device)
    keep = [_f32(t) for t in w8]
    arr = (c_void_p * 8)(*[t.data_ptr() for t in keep])
    _check(load().osd_lat_block_tc(ptr(_f32(x)), ptr(y), arr, ptr(packed), ptr(_f32(film)) if film is not None else c_void_p(0),
                                   ptr(ws), c_int(B), c_int(L), stream()))
    return y


def lat_rmsnorm(x, gamma=None, silu=False):
    B, C = x.shape[0], x.shape[1]
    N = x.numel() // (B * C)
    y = torch.empty_like(x)
    _check(load().osd_lat_rmsnorm(ptr(_f32(x)), ptr(_f32(gamma)) if gamma is not None else c_void_p(0), ptr(y), c_int(B), c_int(C),
                                  ctypes.c_longlong(N), c_int(1 if silu else 0), stream()))
    return y


def lat_conv1x1(x, w, b, act=0, act_channels=0):
    """x [B,Cin,N] (or [B,Cin] for a Linear) , w [Cout,Cin(,1)]"""
    lin = x.dim() == 2
    B, Cin = x.shape[0], x.shape[1]
    N = 1 if lin else x.shape[2]
    Cout = w.shape[0]
    y = torch.empty((B, Cout) if lin else (B, Cout, N), dtype=torch.float32, device=x.device)
    _check(load().osd_lat_conv1x1(ptr(_f32(x)), ptr(_f32(w)), ptr(_f32(b)) if b is not None else c_void_p(0), ptr(y), c_int(B), c_int(Cin),
                                  c_int(Cout), ctypes.c_longlong(N), c_int(act), c_int(act_channels), stream()))
    return y


def lat_conv2d(x, w, b, sh):
    B, Cin, Ain, L = x.shape
    Cout, _, kh, kw = w.shape
    if kw != 3:
        raise OsdError('lat_conv2d: kernel width must be 3')
    Aout = (Ain + 2 - kh) // sh + 1
    y = torch.empty(B, Cout, Aout, L, dtype=torch.float32, device=x.device)
    _check(load().osd_lat_conv2d(ptr(_f32(x)), ptr(_f32(w)), ptr(_f32(b)), ptr(y), c_int(B), c_int(Cin), c_int(Cout), c_int(Ain), c_int(L),
                                 c_int(kh), c_int(sh), stream()))
    return y


def lat_down3(x, w, b):
    B, C, L = x.shape
    y = torch.empty(B, C, L // 3, dtype=torch.float32, device=x.device)
    _check(load().osd_lat_down3(ptr(_f32(x)), ptr(_f32(w)), ptr(_f32(b)), ptr(y), c_int(B), c_int(C), c_int(L), stream()))
    return y


def lat_up3(x, w, b):
    B, C, l = x.shape
    y = torch.empty(B, C, 3 * l, dtype=torch.float32, device=x.device)
    _check(load().osd_lat_up3(ptr(_f32(x)), ptr(_f32(w)), ptr(_f32(b)), ptr(y), c_int(B), c_int(C), c_int(l), stream()))
    return y


def lat_mix(x, p, g):
    """x + p * g; p may have batch 1 (broadcast)"""
    B = x.shape[0]
    per = x.numel() // B
    y = torch.empty_like(x)
    _check(load().osd_lat_mix(ptr(_f32(x)), ptr(_f32(p)), ptr(_f32(g)), ptr(y), c_int(B), ctypes.c_longlong(per), c_int(p.shape[0]), stream()))
    return y
