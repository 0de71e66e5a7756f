"""Drop-in for the reference's pybind extension module `kivi_gemv` (quant/csrc/pybind.cpp:5-8)."""
from __future__ import annotations

import torch

from . import _lib
from .matmul import KIVI_LAYOUT_KERNEL


def _cdiv(a, b):
    return (a + b - 1) // b


def gemv_forward_cuda(_in_feats, _kernel, _scaling_factors, _zeros, bit: int, group_size: int):
    """quant/csrc/gemv_cuda.h:4-10, gemv_cuda.cu:201-246.  in [B, IC] f16, kernel [OC, IC/8] i32,
    scaling_factors/zeros [OC, sf_w] f16 (rows padded as the reference kernels index them, :75,:145) ->
    [B, OC] f16.  4-bit regardless of `bit` (the reference ignores it, :101)."""
    _lib.require_cuda(_in_feats, _kernel, _scaling_factors, _zeros)
    in_feats, kernel = _in_feats.contiguous(), _kernel.contiguous()
    sf, ze = _scaling_factors.contiguous(), _zeros.contiguous()
    Bn, IC = in_feats.shape
    OC = kernel.shape[0]
    if group_size not in (64, 128):
        raise ValueError("gemv_forward_cuda: group_size must be 64 or 128 (the reference launches no kernel "
                         "and returns uninitialised memory otherwise, quant/csrc/gemv_cuda.cu:227-245)")
    out = torch.empty((Bn, OC), dtype=in_feats.dtype, device=in_feats.device)
    with torch.cuda.device(in_feats.device):
        _lib.check(_lib.lib().kivi_gemv_inner_f16(in_feats.data_ptr(), kernel.data_ptr(), sf.data_ptr(), ze.data_ptr(),
                                                   out.data_ptr(), Bn, IC, OC, 4, group_size, sf.shape[1],
                                                   _lib.stream_ptr(in_feats.device)), "gemv_forward_cuda")
    return out


def gemv_forward_cuda_outer_dim(_in_feats, _kernel, _scaling_factors, _zeros, bit: int, group_size: int,
                                nh: int, nh_kv: int):
    """quant/csrc/gemv_cuda.h:12-20, gemv_cuda.cu:511-557.  Kernel-layout operands:
    in [BS, 1, IC] f16, kernel [BSkv, OC/pf, IC] i32, scaling_factors/zeros [BSkv, OC/g, IC] f16
    -> [BS, 1, OC] f16 with OC = zeros.size(1) * group_size (:524)."""
    _lib.require_cuda(_in_feats, _kernel, _scaling_factors, _zeros)
    in_feats, kernel = _in_feats.contiguous(), _kernel.contiguous()
    sf, ze = _scaling_factors.contiguous(), _zeros.contiguous()
    BS, M, IC = in_feats.shape
    OC = ze.shape[1] * group_size
    nh, nh_kv = int(nh), int(nh_kv)
    if nh_kv <= 0:
        raise ValueError("nh_kv must be positive (the reference divides by it, quant/csrc/gemv_cuda.cu:361)")
    if M != 1:
        raise NotImplementedError("only M == 1 is supported (quant/csrc/gemv_cuda.cu:538)")
    out = torch.empty((BS, M, OC), dtype=in_feats.dtype, device=in_feats.device)
    B = BS // nh
    with torch.cuda.device(in_feats.device):
        _lib.check(_lib.lib().kivi_bgemv_outer_f16(
            in_feats.data_ptr(), IC, kernel.data_ptr(), kernel.stride(0), kernel.stride(1),
            sf.data_ptr(), ze.data_ptr(), sf.stride(0), sf.stride(1), out.data_ptr(),
            B, nh, nh_kv, IC, OC, 4 if bit == 4 else 2, group_size, KIVI_LAYOUT_KERNEL,
            _lib.stream_ptr(in_feats.device)), "gemv_forward_cuda_outer_dim")
    return out
