"""numpy face of the C oracle + the restated attention hook.  TEST INFRASTRUCTURE, NOT PRODUCT.

Every function cites the reference file:line it follows (jy-yuan/KIVI @ 876b4d2).
fp16 tensors are numpy float16 arrays; integer codes are int32 arrays.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

from . import build as _build

_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = _build.SO if os.path.exists(_build.SO) and \
            os.path.getmtime(_build.SO) >= os.path.getmtime(_build.SRC) else _build.build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def _f16(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    assert a.dtype == np.float16, a.dtype
    return a


def _chk(rc: int, what: str):
    if rc != 0:
        raise ValueError(f"oracle {what} failed with code {rc}")


# ---------------------------------------------------------------------------------------------
# pack / dequant  (quant/new_pack.py:217-252, :51-83)
# ---------------------------------------------------------------------------------------------
def pack_lastdim(x: np.ndarray, group_size: int, bits: int):
    """triton_quantize_and_pack_along_last_dim (quant/new_pack.py:217-252) on any [..., T] fp16 array."""
    x = _f16(x)
    T = x.shape[-1]
    rows = x.size // T
    fpi = 32 // bits
    assert T % group_size == 0, "quant/new_pack.py:222"
    code = np.zeros(x.shape[:-1] + (T // fpi,), np.int32)
    scale = np.zeros(x.shape[:-1] + (T // group_size,), np.float16)
    mn = np.zeros_like(scale)
    _chk(lib().ko_pack_lastdim_f16(_p(x), ctypes.c_int64(rows), ctypes.c_int64(T), group_size, bits,
                                   _p(code), _p(scale), _p(mn)), "pack_lastdim")
    return code, scale, mn


def unpack_dequant_lastdim(code: np.ndarray, scale: np.ndarray, mn: np.ndarray, group_size: int, bits: int):
    """unpack_and_dequant_vcache (quant/new_pack.py:69-83): fp16 code*scale+mn along the last dim."""
    code = np.ascontiguousarray(code, np.int32)
    fpi = 32 // bits
    T = code.shape[-1] * fpi
    rows = code.size // code.shape[-1]
    out = np.zeros(code.shape[:-1] + (T,), np.float16)
    _chk(lib().ko_unpack_dequant_lastdim_f16(_p(code), _p(_f16(scale)), _p(_f16(mn)), ctypes.c_int64(rows),
                                             ctypes.c_int64(T), group_size, bits, _p(out)), "unpack")
    return out


def unpack_codes_lastdim(code: np.ndarray, bits: int) -> np.ndarray:
    """unpack_tensor (quant/new_pack.py:110-129) along the last dim."""
    fpi = 32 // bits
    w = code.astype(np.uint32)[..., None] >> (np.arange(fpi, dtype=np.uint32) * bits)
    return (w & ((1 << bits) - 1)).astype(np.int32).reshape(code.shape[:-1] + (code.shape[-1] * fpi,))


# ---------------------------------------------------------------------------------------------
# GEMVs
# ---------------------------------------------------------------------------------------------
def bgemv_outer_kernel_layout(inp, w, s, z, bit: int, group_size: int, nh: int, nh_kv: int):
    """kivi_gemv.gemv_forward_cuda_outer_dim (quant/csrc/gemv_cuda.cu:511-557): kernel-layout operands
    inp [BS,1,IC] f16, w [BSkv,OC/pf,IC] i32, s/z [BSkv,OC/g,IC] f16 -> [BS,1,OC] f16."""
    inp = _f16(inp)
    BS, M, IC = inp.shape
    assert M == 1
    w = np.ascontiguousarray(w, np.int32)
    s, z = _f16(s), _f16(z)
    OC = z.shape[1] * group_size                                   # gemv_cuda.cu:524
    out = np.zeros((BS, 1, OC), np.float16)
    _chk(lib().ko_bgemv_outer_kernel_layout(_p(inp), _p(w), _p(s), _p(z), _p(out), BS, IC, OC, bit,
                                            group_size, nh, nh_kv), "bgemv_outer")
    return out


def bmm_fA_qB_outer(group_size: int, fA, qB, scales, zeros, bits: int):
    """cuda_bmm_fA_qB_outer (quant/matmul.py:178-219): fA [B,nh,1,K], qB [B,nh_kv,K,N/fpi],
    scales/zeros [B,nh_kv,K,N/g] -> [B,nh,1,N] fp16."""
    fA = _f16(fA)
    B, nh, M, K = fA.shape
    assert M == 1
    nh_kv = qB.shape[1]
    fpi = 32 // bits
    N = qB.shape[-1] * fpi
    qB = np.ascontiguousarray(qB, np.int32)
    out = np.zeros((B, nh, 1, N), np.float16)
    _chk(lib().ko_bmm_fA_qB_outer(_p(fA), _p(qB), _p(_f16(scales)), _p(_f16(zeros)), _p(out),
                                  B, nh, nh_kv, K, N, bits, group_size), "bmm_fA_qB_outer")
    return out


def gemv_inner_w4(inp, w, s, z, group_size: int):
    """kivi_gemv.gemv_forward_cuda (quant/csrc/gemv_cuda.cu:201-246)."""
    inp = _f16(inp)
    Bn, IC = inp.shape
    w = np.ascontiguousarray(w, np.int32)
    OC = w.shape[0]
    out = np.zeros((Bn, OC), np.float16)
    _chk(lib().ko_gemv_inner_w4(_p(inp), _p(w), _p(_f16(s)), _p(_f16(z)), _p(out), Bn, IC, OC, group_size),
         "gemv_inner_w4")
    return out


def residual_qk(q, k_full):
    """torch.matmul(q, repeat_kv(K_full).transpose(2,3)) on fp16 (models/llama_kivi.py:337): [B,H,1,r]."""
    q, k_full = _f16(q), _f16(k_full)
    B, H, _, D = q.shape
    Hkv, L = k_full.shape[1], k_full.shape[2]
    out = np.zeros((B, H, 1, L), np.float16)
    _chk(lib().ko_residual_gemv_f16(_p(q), _p(k_full), _p(out), B, H, Hkv, L, D, 0, ctypes.c_int64(D)), "res_qk")
    return out


def residual_pv(p, v_full):
    """torch.matmul(p[..., -L:], repeat_kv(V_full)) on fp16 (models/llama_kivi.py:384): [B,H,1,D]."""
    p, v_full = _f16(p), _f16(v_full)
    B, H, _, L = p.shape
    Hkv, L2, D = v_full.shape[1], v_full.shape[2], v_full.shape[3]
    assert L == L2
    out = np.zeros((B, H, 1, D), np.float16)
    _chk(lib().ko_residual_gemv_f16(_p(p), _p(v_full), _p(out), B, H, Hkv, L, D, 1, ctypes.c_int64(L)), "res_pv")
    return out


def scale_softmax(logits, head_dim: int, mask=None):
    """/ sqrt(head_dim) in fp16, optional mask, fp32 softmax -> fp16 (models/llama_kivi.py:339,369-375)."""
    logits = _f16(logits)
    T = logits.shape[-1]
    rows = logits.size // T
    probs = np.zeros_like(logits)
    m = None
    if mask is not None:
        m = _f16(np.broadcast_to(mask, logits.shape))
    _chk(lib().ko_scale_softmax_f16(_p(logits), _p(m) if m is not None else None, _p(probs),
                                    ctypes.c_int64(rows), ctypes.c_int64(T),
                                    ctypes.c_float(math.sqrt(head_dim))), "softmax")
    return probs


def add_f16(a, b):
    a, b = _f16(a), _f16(b)
    out = np.zeros_like(a)
    _chk(lib().ko_add_f16(_p(a), _p(b), _p(out), ctypes.c_int64(a.size)), "add")
    return out


# ---------------------------------------------------------------------------------------------
# The attention hook, restated (models/llama_kivi.py:314-399 decode, :425-455 prefill split)
# cache = (Kq_code[B,Hkv,D,Tk/fpi] i32 | None, K_full[B,Hkv,r,D] f16 | None, K_scale, K_mn,
#          Vq_code[B,Hkv,Tv,D/fpi] i32 | None, V_full[B,Hkv,L,D] f16, V_scale, V_mn, kv_seq_len)
# ---------------------------------------------------------------------------------------------
def prefill_cache(k, v, group_size: int, k_bits: int, v_bits: int, residual_length: int):
    """Split + quantise the prompt's K/V exactly as models/llama_kivi.py:425-455."""
    k, v = _f16(k), _f16(v)
    n = k.shape[-2]
    R = residual_length
    if n % R != 0:                                                 # :425-431
        if n < R:
            kq, kfull = None, k
        else:
            kq, kfull = k[:, :, :-(n % R), :], np.ascontiguousarray(k[:, :, -(n % R):, :])
    else:                                                          # :432-434
        kq, kfull = k, None
    if kq is not None:                                             # :435-436
        Kq, Ks, Kz = pack_lastdim(np.ascontiguousarray(kq.transpose(0, 1, 3, 2)), group_size, k_bits)
    else:
        Kq = Ks = Kz = None
    if n <= R:                                                     # :442-446
        Vq = Vs = Vz = None
        vfull = v
    else:                                                          # :447-452
        Vq, Vs, Vz = pack_lastdim(np.ascontiguousarray(v[:, :, :-R, :]), group_size, v_bits)
        vfull = np.ascontiguousarray(v[:, :, -R:, :])
    return (Kq, kfull, Ks, Kz, Vq, vfull, Vs, Vz, n)


def decode_step(cache, q, k_new, v_new, group_size: int, k_bits: int, v_bits: int, residual_length: int,
                mask=None):
    """One decode step of LlamaFlashAttention_KIVI.forward (models/llama_kivi.py:314-399).
    q [B,H,1,D], k_new/v_new [B,Hkv,1,D] (post-RoPE).  Returns (attn_output [B,H,1,D] f16, probs, new cache)."""
    Kq, Kfull, Ks, Kz, Vq, Vfull, Vs, Vz, kv_len = cache
    q, k_new, v_new = _f16(q), _f16(k_new), _f16(v_new)
    D = q.shape[-1]
    R = residual_length
    kv_len = kv_len + 1
    att_q = bmm_fA_qB_outer(group_size, q, Kq, Ks, Kz, k_bits) if Kq is not None else None   # :323-325
    Kfull = np.concatenate([Kfull, k_new], axis=2) if Kfull is not None else k_new          # :333-336
    att_f = residual_qk(q, Kfull)                                                            # :337
    logits = np.concatenate([att_q, att_f], axis=-1) if att_q is not None else att_f        # :338-341
    if Kfull.shape[-2] == R:                                                                 # :343-356
        assert R % group_size == 0
        nq, ns, nz = pack_lastdim(np.ascontiguousarray(Kfull.transpose(0, 1, 3, 2)), group_size, k_bits)
        Kfull = None
        if Kq is not None:
            Kq = np.concatenate([Kq, nq], axis=3)
            Ks = np.concatenate([Ks, ns], axis=3)
            Kz = np.concatenate([Kz, nz], axis=3)
        else:
            Kq, Ks, Kz = nq, ns, nz
    assert logits.shape[-1] == kv_len                                                        # :358-362
    probs = scale_softmax(logits, D, mask)                                                   # :339,364-375
    Vfull = np.concatenate([Vfull, v_new], axis=2)                                           # :377
    L = Vfull.shape[-2]
    if Vq is None:
        out = residual_pv(probs, Vfull)                                                      # :380
    else:
        out = bmm_fA_qB_outer(group_size, np.ascontiguousarray(probs[..., :-L]), Vq, Vs, Vz, v_bits)  # :382-383
        out = add_f16(out, residual_pv(np.ascontiguousarray(probs[..., -L:]), Vfull))        # :384
    if L > R:                                                                                # :386-399
        assert L == R + 1
        nq, ns, nz = pack_lastdim(np.ascontiguousarray(Vfull[:, :, :1, :]), group_size, v_bits)
        Vfull = np.ascontiguousarray(Vfull[:, :, 1:, :])
        if Vq is not None:
            Vq = np.concatenate([Vq, nq], axis=2)
            Vs = np.concatenate([Vs, ns], axis=2)
            Vz = np.concatenate([Vz, nz], axis=2)
        else:
            Vq, Vs, Vz = nq, ns, nz
    return out, probs, (Kq, Kfull, Ks, Kz, Vq, Vfull, Vs, Vz, kv_len)
