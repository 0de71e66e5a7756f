"""Drop-in surface of the importable helpers of the reference's quant/gemv.py (a test/bench script)."""
from __future__ import annotations

import torch

from . import _lib, kivi_gemv  # noqa: F401  (re-exported like `import kivi_gemv` in quant/gemv.py:12)
from .new_pack import pack_tensor  # noqa: F401  (quant/gemv.py:10)


def dequant_weight(w, scale, mn, gs):
    """quant/gemv.py:64-67 (fp16 dequant along the inner dim; the reference tests' comparison oracle)."""
    w_fp = w.half().view(w.shape[0], w.shape[1] // gs, gs)
    w_fp = w_fp * scale.unsqueeze(-1) + mn.unsqueeze(-1)
    return w_fp.view(w.shape)


def dequant_weight_outer(w, scale, mn, gs):
    """quant/gemv.py:70-74 (fp16 dequant along the outer dim)."""
    w_fp = w.half().view(w.shape[0], w.shape[1], w.shape[2] // gs, gs)
    w_fp = w_fp * scale.unsqueeze(-1) + mn.unsqueeze(-1)
    return w_fp.view(w.shape)


def gemv_fwd(bit, group_size, inp, qweight, mn, scale):
    """quant/gemv.py:77-90 (Triton gemv_kernel_g64 :16-61): inp [B, IC] f16, qweight [OC, IC/pf] i32,
    mn/scale [OC, IC/g] f16 (unpadded) -> [B, OC] f16.  The reference asserts group_size == 64 (:83)
    because its Triton kernel hard-codes it; any group_size % pack_factor == 0 works here."""
    _lib.require_cuda(inp, qweight, mn, scale)
    B, IC = inp.shape
    OC = qweight.shape[0]
    inp, qweight, mn, scale = inp.contiguous(), qweight.contiguous(), mn.contiguous(), scale.contiguous()
    output = torch.empty((B, OC), device=inp.device, dtype=torch.float16)
    with torch.cuda.device(inp.device):
        _lib.check(_lib.lib().kivi_gemv_inner_f16(inp.data_ptr(), qweight.data_ptr(), scale.data_ptr(), mn.data_ptr(),
                                                   output.data_ptr(), B, IC, OC, bit, group_size, scale.shape[1],
                                                   _lib.stream_ptr(inp.device)), "gemv_fwd")
    return output
