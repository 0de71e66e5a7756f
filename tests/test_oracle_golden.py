"""Pin the CPU oracle against golden vectors produced by the REFERENCE's own Python
(tests/golden/make_golden.py: quant/new_pack.py pure-torch helpers, models/utils_quant.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import fake_quant, ref


@pytest.fixture(scope="module")
def pack_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "pack_reference.npz"))


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("g", [32, 64])
def test_pack_v_matches_reference(pack_gold, bits, g):
    """oracle pack == quant_and_pack_vcache (quant/new_pack.py:30-48) bit for bit."""
    tag = f"b{bits}_g{g}"
    code, scale, mn = ref.pack_lastdim(pack_gold[f"v_{tag}"], g, bits)
    np.testing.assert_array_equal(code, pack_gold[f"v_code_{tag}"])
    np.testing.assert_array_equal(scale.view(np.uint16), pack_gold[f"v_scale_{tag}"].view(np.uint16))
    np.testing.assert_array_equal(mn.view(np.uint16), pack_gold[f"v_mn_{tag}"].view(np.uint16))


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("g", [32, 64])
def test_pack_k_matches_reference(pack_gold, bits, g):
    """oracle pack of K^T == transpose of quant_and_pack_kcache (quant/new_pack.py:8-27)."""
    tag = f"b{bits}_g{g}"
    k = pack_gold[f"k_{tag}"]                                        # [B,nh,T,D]
    code, scale, mn = ref.pack_lastdim(np.ascontiguousarray(k.transpose(0, 1, 3, 2)), g, bits)
    np.testing.assert_array_equal(code, pack_gold[f"k_code_{tag}"].transpose(0, 1, 3, 2))
    np.testing.assert_array_equal(scale.view(np.uint16),
                                  np.ascontiguousarray(pack_gold[f"k_scale_{tag}"].transpose(0, 1, 3, 2)).view(np.uint16))
    np.testing.assert_array_equal(mn.view(np.uint16),
                                  np.ascontiguousarray(pack_gold[f"k_mn_{tag}"].transpose(0, 1, 3, 2)).view(np.uint16))


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("g", [32, 64])
def test_dequant_matches_reference(pack_gold, bits, g):
    """oracle dequant == unpack_and_dequant_vcache / _kcache (quant/new_pack.py:51-83)."""
    tag = f"b{bits}_g{g}"
    deq = ref.unpack_dequant_lastdim(pack_gold[f"v_code_{tag}"], pack_gold[f"v_scale_{tag}"],
                                     pack_gold[f"v_mn_{tag}"], g, bits)
    np.testing.assert_array_equal(deq.view(np.uint16), pack_gold[f"v_deq_{tag}"].view(np.uint16))
    kc = np.ascontiguousarray(pack_gold[f"k_code_{tag}"].transpose(0, 1, 3, 2))
    ks = np.ascontiguousarray(pack_gold[f"k_scale_{tag}"].transpose(0, 1, 3, 2))
    km = np.ascontiguousarray(pack_gold[f"k_mn_{tag}"].transpose(0, 1, 3, 2))
    kdeq = ref.unpack_dequant_lastdim(kc, ks, km, g, bits).transpose(0, 1, 3, 2)
    np.testing.assert_array_equal(np.ascontiguousarray(kdeq).view(np.uint16), pack_gold[f"k_deq_{tag}"].view(np.uint16))


def test_unpack_codes_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "pack_tensor_reference.npz"))
    for bits in (2, 4, 8):
        np.testing.assert_array_equal(ref.unpack_codes_lastdim(gold[f"pack_d3_b{bits}"], bits), gold[f"data_b{bits}"])


@pytest.mark.parametrize("bits", [2, 4])
def test_fake_quant_port_matches_reference(golden_dir, bits):
    """oracle/fake_quant.py == models/utils_quant.py (simulate paths) bit for bit on CPU."""
    gold = np.load(os.path.join(golden_dir, "fake_quant_reference.npz"))
    k = torch.from_numpy(gold[f"k_b{bits}"])
    v = torch.from_numpy(gold[f"v_b{bits}"])
    g = 32
    codes, sc, mn = fake_quant.quantize_by_channel_and_pack_cache_sim(k.clone(), g, bits)
    np.testing.assert_array_equal(codes.numpy(), gold[f"k_codes_b{bits}"])
    kdq = fake_quant.dequantize_by_channel_and_unpack_cache_sim(codes, g, k.shape, bits, sc, mn)
    np.testing.assert_array_equal(kdq.numpy().view(np.uint16), gold[f"k_fake_b{bits}"].view(np.uint16))
    B, H, T, D = v.shape
    vdq = fake_quant.asym_grouped_quantizer(v.transpose(1, 2).reshape(B, T, H * D).clone(), bits, g)
    vdq = vdq.view(B, T, H, D).transpose(1, 2)
    np.testing.assert_array_equal(vdq.numpy().view(np.uint16), gold[f"v_fake_b{bits}"].view(np.uint16))


@pytest.mark.parametrize("bits", [2, 4])
def test_fake_quant_equals_packed_dequant(golden_dir, bits):
    """Cross-check: the reference fake-quant K (g-token groups per channel) equals
    dequant(pack) of the packed-cache path, and fake-quant V equals dequant(pack) per token."""
    gold = np.load(os.path.join(golden_dir, "fake_quant_reference.npz"))
    k, v = gold[f"k_b{bits}"], gold[f"v_b{bits}"]
    g = 32
    code, scale, mn = ref.pack_lastdim(np.ascontiguousarray(k.transpose(0, 1, 3, 2)), g, bits)
    kdeq = ref.unpack_dequant_lastdim(code, scale, mn, g, bits).transpose(0, 1, 3, 2)
    np.testing.assert_array_equal(np.ascontiguousarray(kdeq).view(np.uint16), gold[f"k_fake_b{bits}"].view(np.uint16))
    code, scale, mn = ref.pack_lastdim(v, g, bits)
    vdeq = ref.unpack_dequant_lastdim(code, scale, mn, g, bits)
    np.testing.assert_array_equal(vdeq.view(np.uint16), gold[f"v_fake_b{bits}"].view(np.uint16))
