"""torch-CPU restatement of the reference's fake-quant path (models/utils_quant.py).

TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT.  This is the "reference's own CPU
implementation of the path" that bench.py times as `cpu_baseline` (kind "port") and as the
`--impl reference` arm: /root/reference cannot travel to the GPU box, so the functions it would
run are restated here op-for-op (same torch calls, same order) and pinned bit-for-bit against
golden vectors produced by the reference itself (tests/golden/fake_quant_reference.npz,
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def process_input_by_channel(inp: torch.Tensor, group_size: int):
    """models/utils_quant.py:418-432"""
    num_features = inp.shape[-1]
    input_flatten = inp.view(-1, num_features).transpose(0, 1)
    num_instances = input_flatten.shape[-1]
    if num_instances % group_size != 0:
        new_num_instances = (num_instances // group_size + 1) * group_size
        delta = new_num_instances - num_instances
        input_flatten = torch.cat(
            [input_flatten, torch.zeros([num_features, delta], dtype=inp.dtype, device=inp.device)], 1)
    input_groups = input_flatten.reshape(-1, group_size)
    mn, mx = torch.min(input_groups, 1)[0], torch.max(input_groups, 1)[0]
    return input_groups.view(num_features, -1, group_size), mn.view(num_features, -1), mx.view(num_features, -1)


def quantize_by_channel_and_pack_cache_sim(inp: torch.Tensor, group_size: int, num_bits: int):
    """models/utils_quant.py:498-521, simulate=True branch: returns (rounded codes as float, scale, mn)."""
    assert len(inp.shape) == 4
    bsz, _, seq_len, _ = inp.shape
    inp = inp.transpose(1, 2).reshape(bsz, seq_len, -1)
    input_groups, mn, mx = process_input_by_channel(inp, group_size)
    mn, mx = mn.unsqueeze(-1), mx.unsqueeze(-1)
    scale = (mx - mn) / (2 ** num_bits - 1)
    input_groups = (input_groups - mn) / scale
    input_groups = F.relu(input_groups)
    rounded_input = input_groups.round_()
    return rounded_input, scale, mn


def dequantize_by_channel_and_unpack_cache_sim(data, group_size, shape, bits, scale, mn):
    """models/utils_quant.py:533-563, simulate=True branch."""
    assert len(shape) == 4
    num_feats = shape[1] * shape[3]
    ori_num_instances = shape[0] * shape[2]
    data = data * scale + mn
    dequantized_input = data.view(num_feats, -1)
    if ori_num_instances != dequantized_input.shape[1]:
        dequantized_input = dequantized_input[:, 0:ori_num_instances]
    data = dequantized_input.transpose(0, 1).view(shape[0], -1, num_feats)
    data = data.view(shape[0], shape[2], shape[1], -1).transpose(1, 2)
    assert data.shape == shape
    return data


def asym_grouped_quantizer(inp: torch.Tensor, num_bits: int, group_size: int):
    """AsymGroupedQuantizer.forward, models/utils_quant.py:167-206 (input [bs, seqlen, d])."""
    bs, seqlen, d = inp.shape
    num_groups = d // group_size
    if num_groups * group_size != inp.shape[-1]:
        raise ValueError("group_size should be a factor of the last dimension size")
    input_in_groups = inp.view(bs, seqlen, num_groups, group_size)
    mx, mn = input_in_groups.max(dim=-1)[0], input_in_groups.min(dim=-1)[0]
    mx, mn = mx.unsqueeze(-1), mn.unsqueeze(-1)
    scale = (mx - mn) / (2 ** num_bits - 1)
    input_in_groups = (input_in_groups - mn) / scale
    input_in_groups = F.relu(input_in_groups)
    rounded_input_in_groups = input_in_groups.round_()
    dequantized_input_in_groups = rounded_input_in_groups * scale + mn
    return dequantized_input_in_groups.view(bs, seqlen, -1)


def fake_quant_decode_attention(q, k, v, group_size: int, k_bits: int, v_bits: int, compute_dtype=torch.float32):
    """One attention layer of fake-quant decode, the CPU baseline of BASELINE.md section 2:
    per-channel K in g-token groups (quantize->dequantize), per-token V, then softmax(qK^T/sqrt(D)) V.
    q [B,H,1,D], k/v [B,Hkv,T,D] (T % group_size == 0 as in the packed-cache path)."""
    B, Hkv, T, D = k.shape
    H = q.shape[1]
    codes, sc, mn = quantize_by_channel_and_pack_cache_sim(k, group_size, k_bits)
    k_fake = dequantize_by_channel_and_unpack_cache_sim(codes, group_size, k.shape, k_bits, sc, mn)
    v3 = v.transpose(1, 2).reshape(B, T, Hkv * D)
    v_fake = asym_grouped_quantizer(v3, v_bits, group_size).view(B, T, Hkv, D).transpose(1, 2)
    rep = H // Hkv
    if rep > 1:
        k_fake = k_fake.repeat_interleave(rep, dim=1)
        v_fake = v_fake.repeat_interleave(rep, dim=1)
    qf, kf, vf = q.to(compute_dtype), k_fake.to(compute_dtype), v_fake.to(compute_dtype)
    att = torch.matmul(qf, kf.transpose(2, 3)) / math.sqrt(D)
    att = torch.softmax(att, dim=-1, dtype=torch.float32).to(compute_dtype)
    return torch.matmul(att, vf)
