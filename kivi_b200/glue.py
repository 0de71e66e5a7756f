"""ctypes wrappers of the decode-step glue kernels (csrc/kivi_model.cu): residual-add + RMSNorm,
RoPE + q/k/v split, SiLU*mul.  fp16 CUDA tensors only."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_B = False


def _bind():
    global _B
    if _B:
        return
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    _lib.bind("kivi_add_rmsnorm_f16", i32, [vp, vp, vp, vp, i32, i32, ctypes.c_float, vp])
    _lib.bind("kivi_rope_split_f16", i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp])
    _lib.bind("kivi_silu_mul_f16", i32, [vp, vp, i32, i32, vp])
    _lib.bind("kivi_greedy_sample_exchange_f32", i32, [vp, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp])
    _B = True


def add_rmsnorm(x, residual, weight, out, eps: float):
    """residual += x (x may be None); out = weight * fp16(residual * rsqrt(mean(residual^2) + eps))."""
    _bind()
    rows, hidden = residual.shape
    _lib.check(_lib.lib().kivi_add_rmsnorm_f16(x.data_ptr() if x is not None else None, residual.data_ptr(),
                                               weight.data_ptr(), out.data_ptr(), rows, hidden, eps,
                                               _lib.stream_ptr(residual.device)), "kivi_add_rmsnorm_f16")
    return out


def rope_split(qkv, cos_table, sin_table, pos, q, k, v):
    """qkv [B,(H+2Hkv)*128] -> q [B,H,128], k [B,Hkv,128] rotated at position pos[b] (int64), v [B,Hkv,128]."""
    _bind()
    B, H, Hkv = q.shape[0], q.shape[1], k.shape[1]
    _lib.check(_lib.lib().kivi_rope_split_f16(qkv.data_ptr(), cos_table.data_ptr(), sin_table.data_ptr(), pos.data_ptr(),
                                              q.data_ptr(), k.data_ptr(), v.data_ptr(), B, H, Hkv, cos_table.shape[0],
                                              _lib.stream_ptr(qkv.device)), "kivi_rope_split_f16")


def silu_mul(gate_up, out):
    _bind()
    rows, inter = out.shape
    _lib.check(_lib.lib().kivi_silu_mul_f16(gate_up.data_ptr(), out.data_ptr(), rows, inter,
                                            _lib.stream_ptr(out.device)), "kivi_silu_mul_f16")
    return out


def greedy_sample(logits, next_local, ids_feedback=None, exchange=None):
    """next_local[b] = argmax(logits[b]) (+ copy into ids_feedback); with `exchange` (kivi_b200.dist.PeerTokenExchange) the
    same kernel also stores the ids into every rank's token buffer over NVLink and waits for the other ranks' ids."""
    _bind()
    B, V = logits.shape
    assert logits.dtype == torch.float32 and logits.is_contiguous() and next_local.dtype == torch.int64
    ex = exchange
    _lib.check(_lib.lib().kivi_greedy_sample_exchange_f32(
        logits.data_ptr(), B, V, next_local.data_ptr(), ids_feedback.data_ptr() if ids_feedback is not None else None,
        ex.peer_ptrs.data_ptr() if ex is not None else None, ex.rank if ex is not None else 0, ex.world if ex is not None else 1,
        ex.step.data_ptr() if ex is not None else None, ex.err.data_ptr() if ex is not None else None,
        _lib.stream_ptr(logits.device)), "kivi_greedy_sample_exchange_f32")
    return next_local
