"""GPU-side restatement of the reference's pack for FULL-SIZE parity cases.  TEST INFRASTRUCTURE, NOT PRODUCT.

The C oracle (kivi_oracle.c) is single-threaded and checks small cases; the reference's own pinned cases
(quant/test.py:21-54: 555 x 32 x 433 x 128 and 11 x 32 x 4096 x 128 elements) need a checker that runs where the
reference would: on the GPU, with the reference's own ATen ops.  `triton_quantize_and_pack_along_last_dim`
(quant/new_pack.py:217-252) is two Triton kernels doing exact operations (group min/max :158-177, OR-pack :132-154)
around an ATen elementwise chain (:238-242); the chain below is that chain op for op, the two Triton kernels are
replaced by the equivalent exact torch ops.  Pinned in tests/test_pack_gpu.py against the C oracle (itself pinned to
the reference's Python by tests/golden/pack_reference.npz).  Nothing under kivi_b200/ imports this module.
"""
from __future__ import annotations

import torch


def pack_lastdim(data: torch.Tensor, group_size: int, bit: int, chunk_rows: int = 1 << 22):
    """quant/new_pack.py:217-252 on a [..., T] fp16 CUDA tensor -> (code int32 [..., T/fpi], scale, mn [..., T/g])."""
    assert data.dtype == torch.float16 and data.is_cuda
    shape = data.shape
    T = shape[-1]
    assert T % group_size == 0                                           # :222
    fpi = 32 // bit
    flat = data.reshape(-1, T // group_size, group_size)                # :227
    rows = flat.shape[0]
    code = torch.empty((rows, T // fpi), dtype=torch.int32, device=data.device)
    scale = torch.empty((rows, T // group_size), dtype=torch.float16, device=data.device)
    mn = torch.empty_like(scale)
    shifts = (torch.arange(fpi, device=data.device, dtype=torch.int32) * bit)
    for lo in range(0, rows, chunk_rows):                                # bounded temporaries (the int32 codes are 2x the input)
        x = flat[lo:lo + chunk_rows]
        mnc = x.amin(-1)                                                 # _minmax_along_last_dim :158-177 (exact)
        mxc = x.amax(-1)
        sc = (mxc - mnc) / (2 ** bit - 1)                                # :238
        d = x - mnc.unsqueeze(-1)                                        # :239
        d.div_(sc.unsqueeze(-1))                                         # :240
        d = d.clamp_(0, 2 ** bit - 1).round_().to(torch.int32)           # :241  (NaN of a flat group -> 0 on CUDA)
        d = d.view(d.shape[0], T // fpi, fpi)
        code[lo:lo + chunk_rows] = (d << shifts).sum(-1, dtype=torch.int32)   # _pack_along_last_dim :132-154: OR of disjoint fields
        scale[lo:lo + chunk_rows] = sc
        mn[lo:lo + chunk_rows] = mnc
    return (code.view(shape[:-1] + (T // fpi,)), scale.view(shape[:-1] + (T // group_size,)),
            mn.view(shape[:-1] + (T // group_size,)))


def unpack_dequant_lastdim(code: torch.Tensor, scale: torch.Tensor, mn: torch.Tensor, group_size: int, bits: int):
    """unpack_and_dequant_vcache (quant/new_pack.py:69-83) with scale / mn already [..., T/g]: fp16 code*scale + mn."""
    fpi = 32 // bits
    shifts = (torch.arange(fpi, device=code.device, dtype=torch.int32) * bits)
    c = ((code.unsqueeze(-1) >> shifts) & (2 ** bits - 1)).to(torch.float16)          # unpack_tensor :110-129
    c = c.view(code.shape[:-1] + (code.shape[-1] * fpi // group_size, group_size))
    out = c * scale.unsqueeze(-1) + mn.unsqueeze(-1)                                  # :81-82 (fp16 mul, fp16 add)
    return out.view(code.shape[:-1] + (code.shape[-1] * fpi,))
