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
LIB_PATH = os.path.join(_HERE, 'libosd_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'osd_b200.h')

_lib = None

BF16, TF32 = 0, 1
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
    return c_void_p(t.data_ptr())


def stream():
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


def rope_table(L: int, device) -> torch.Tensor:
    """[L, 2, 32] fp32 cos|sin table; inv_freq formed exactly as osu_dreamer/common/attn.py:16-18."""
    inv_freq = (10000 ** (torch.arange(0, 64, 2).float() / -64)).contiguous()
    arr = (c_float * 32)(*inv_freq.tolist())
    out = torch.empty(L, 2, 32, dtype=torch.float32, device=device)
    _check(load().osd_rope_table(arr, c_int(L), ptr(out), stream()))
    return out


def qkv_proj(x, w, bias, qnorm_w, knorm_w, rope, L, raw_out=None):
    T = x.shape[0]
    out = torch.empty(T, 3072, dtype=torch.bfloat16, device=x.device)
    _check(load().osd_qkv_proj(ptr(x), ptr(w), ptr(bias), ptr(qnorm_w), ptr(knorm_w), ptr(rope), ptr(out),
                               ptr(raw_out), c_int(T), c_int(L), c_int(_elem_of(x)), stream()))
    return out
