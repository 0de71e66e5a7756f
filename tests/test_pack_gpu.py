"""Parity of the CUDA pack path (kivi_pack.cu through the C ABI / Python surface) with the oracle and
with golden vectors of the reference: BIT-EXACT on codes, scale and mn."""
import os

import numpy as np
import pytest
import torch

from oracle import ref
from tests._util import to_np

pytestmark = pytest.mark.gpu


def _pack_gpu(x_np, g, bits):
    from kivi_b200 import new_pack
    x = torch.from_numpy(x_np).cuda()
    code, scale, mn = new_pack._pack_lastdim(x, g, bits)
    torch.cuda.synchronize()
    return to_np(code), to_np(scale), to_np(mn)


def _assert_pack_equal(x, g, bits):
    code, scale, mn = _pack_gpu(x, g, bits)
    rc, rs, rm = ref.pack_lastdim(x, g, bits)
    np.testing.assert_array_equal(code, rc)
    np.testing.assert_array_equal(scale.view(np.uint16), rs.view(np.uint16))
    np.testing.assert_array_equal(mn.view(np.uint16), rm.view(np.uint16))


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("g", [32, 64, 128])
@pytest.mark.parametrize("shape", [(2, 4, 1, 128), (2, 3, 128, 128), (1, 2, 128, 256), (3, 1, 7, 384)])
def test_pack_matches_oracle(bits, g, shape):
    """Decode V token [B,Hkv,1,D], K flush [B,Hkv,D,R], prefill-like shapes."""
    rng = np.random.default_rng(hash((bits, g, shape)) % 2**32)
    x = (rng.standard_normal(shape) * rng.uniform(0.1, 4.0) + rng.uniform(-1, 1)).astype(np.float16)
    _assert_pack_equal(x, g, bits)


@pytest.mark.parametrize("bits,g", [(2, 16), (2, 48), (4, 8), (4, 24), (8, 4), (2, 8), (4, 4), (8, 12), (2, 512), (4, 512)])
def test_pack_unusual_group_sizes(bits, g):
    """Group sizes outside the fast path (lanes-per-group not a power of two, g < fpi, g = 512...)."""
    rng = np.random.default_rng(bits * 100 + g)
    T = int(np.lcm(np.lcm(g, 32 // bits), 64) * 2)
    x = rng.standard_normal((3, 5, T)).astype(np.float16)
    _assert_pack_equal(x, g, bits)


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_pack_edge_values(bits):
    """Degenerate groups (mx == mn -> code 0, scale 0; the reference's CPU path corrupts the word here,
    SURVEY 8 a1), ties at .5, fp16 subnormals, large magnitudes, exact grid points."""
    g = 32
    rng = np.random.default_rng(5)
    rows = []
    rows.append(np.full(64, 1.25))                                  # constant -> degenerate
    rows.append(np.zeros(64))
    rows.append(np.concatenate([np.linspace(0, 3, 32), np.linspace(-7, 8, 32)]))      # ties / grid points
    rows.append(rng.standard_normal(64) * 6e-6)                     # subnormal range
    rows.append(rng.standard_normal(64) * 2e4)                      # large
    rows.append(np.concatenate([[-60000.0, 60000.0], rng.standard_normal(62)]))       # d overflows? (inf scale)
    rows.append(np.arange(64) % (2 ** bits) * 0.5)                  # exact levels
    x = np.stack(rows).astype(np.float16)
    _assert_pack_equal(x, g, bits)


def test_pack_empty_and_errors():
    from kivi_b200 import new_pack
    x = torch.empty((2, 3, 0, 128), dtype=torch.float16, device="cuda")
    code, scale, mn = new_pack.triton_quantize_and_pack_along_last_dim(x, 32, 2)
    assert code.shape == (2, 3, 0, 8) and scale.shape == (2, 3, 0, 4)
    with pytest.raises(AssertionError):                             # quant/new_pack.py:222
        new_pack.triton_quantize_and_pack_along_last_dim(torch.zeros((1, 1, 4, 48), dtype=torch.float16, device="cuda"), 32, 2)
    with pytest.raises(RuntimeError):                               # no CPU fallback
        new_pack.triton_quantize_and_pack_along_last_dim(torch.zeros((1, 1, 4, 64), dtype=torch.float16), 32, 2)


def test_pack_unaligned_view():
    """A storage offset that breaks 16-byte alignment takes the scalar-load path; same bits."""
    rng = np.random.default_rng(9)
    base = torch.from_numpy(rng.standard_normal(4 * 128 + 8).astype(np.float16)).cuda()
    from kivi_b200 import _lib
    x = base[3:3 + 4 * 128].view(4, 128)                            # offset 6 bytes
    assert x.data_ptr() % 16 != 0
    code = torch.empty((4, 8), dtype=torch.int32, device="cuda")
    scale = torch.empty((4, 4), dtype=torch.float16, device="cuda")
    mn = torch.empty_like(scale)
    _lib.check(_lib.lib().kivi_pack_lastdim_f16(x.data_ptr(), 4, 128, 32, 2, code.data_ptr(), scale.data_ptr(),
                                                mn.data_ptr(), _lib.stream_ptr()), "pack")
    rc, rs, rm = ref.pack_lastdim(to_np(x), 32, 2)
    np.testing.assert_array_equal(to_np(code), rc)
    np.testing.assert_array_equal(to_np(scale).view(np.uint16), rs.view(np.uint16))


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("g", [32, 64])
def test_surface_matches_reference_golden(golden_dir, bits, g):
    """quant_and_pack_vcache / _kcache / unpack_and_dequant_* through the drop-in surface == the
    reference's own outputs (tests/golden/pack_reference.npz, made by quant/new_pack.py on CPU)."""
    from kivi_b200 import new_pack
    gold = np.load(os.path.join(golden_dir, "pack_reference.npz"))
    tag = f"b{bits}_g{g}"
    v = torch.from_numpy(gold[f"v_{tag}"]).cuda()
    code, scale, mn = new_pack.quant_and_pack_vcache(v, g, bits)
    np.testing.assert_array_equal(to_np(code), gold[f"v_code_{tag}"])
    np.testing.assert_array_equal(to_np(scale.squeeze(-1)).view(np.uint16), gold[f"v_scale_{tag}"].view(np.uint16))
    np.testing.assert_array_equal(to_np(mn.squeeze(-1)).view(np.uint16), gold[f"v_mn_{tag}"].view(np.uint16))
    deq = new_pack.unpack_and_dequant_vcache(code, scale, mn, g, bits)
    np.testing.assert_array_equal(to_np(deq).view(np.uint16), gold[f"v_deq_{tag}"].view(np.uint16))
    k = torch.from_numpy(gold[f"k_{tag}"]).cuda()
    kcode, kscale, kmn = new_pack.quant_and_pack_kcache(k, g, bits)
    np.testing.assert_array_equal(to_np(kcode), gold[f"k_code_{tag}"])
    np.testing.assert_array_equal(to_np(kscale.squeeze(-2)).view(np.uint16), gold[f"k_scale_{tag}"].view(np.uint16))
    np.testing.assert_array_equal(to_np(kmn.squeeze(-2)).view(np.uint16), gold[f"k_mn_{tag}"].view(np.uint16))
    kdeq = new_pack.unpack_and_dequant_kcache(kcode, kscale, kmn, g, bits)
    np.testing.assert_array_equal(to_np(kdeq).view(np.uint16), gold[f"k_deq_{tag}"].view(np.uint16))


def test_pack_tensor_surface_matches_reference_golden(golden_dir):
    from kivi_b200 import new_pack
    gold = np.load(os.path.join(golden_dir, "pack_tensor_reference.npz"))
    for bits in (2, 4, 8):
        data = torch.from_numpy(gold[f"data_b{bits}"]).cuda()
        for dim in (2, 3):
            packed = new_pack.pack_tensor(data, bits, dim)
            np.testing.assert_array_equal(to_np(packed), gold[f"pack_d{dim}_b{bits}"])
            np.testing.assert_array_equal(to_np(new_pack.unpack_tensor(packed, bits, dim)).astype(np.int32),
                                          gold[f"unpack_d{dim}_b{bits}"])


@pytest.mark.parametrize("bits,g,shape", [(2, 32, (32, 32, 128, 128)), (2, 32, (32, 32, 1, 128)), (4, 64, (16, 8, 128, 64))])
def test_pack_full_size_properties(bits, g, shape):
    """BASELINE sizes (cfg 2 K flush [32,32,128,R=128] = 33.5 MB, decode V [32,32,1,128]; cfg 4 K flush):
    a random slab is compared bit-exactly with the oracle; over the whole tensor: codes in range,
    scale == fp16((mx-mn)/maxq), mn == group min, |x - dequant| <= scale/2 (+1 fp16 ulp)."""
    from kivi_b200 import new_pack
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(shape, generator=gen, device="cuda", dtype=torch.float16)
    code, scale, mn = new_pack.triton_quantize_and_pack_along_last_dim(x, g, bits)
    maxq = 2 ** bits - 1
    xg = x.view(shape[:-1] + (shape[-1] // g, g))
    assert torch.equal(mn, xg.min(-1)[0])
    exp_scale = ((xg.max(-1)[0] - mn) / maxq)
    assert torch.equal(scale, exp_scale)
    deq = new_pack._unpack_dequant_lastdim(code, scale, mn, g, bits)
    err = (deq.float() - x.float()).abs().view(xg.shape)
    # t2 = fp16((x - mn) / scale) is rounded to fp16 before rint(): up to 2^-7 code units near 15 (4-bit)
    bound = scale.float().unsqueeze(-1) * (0.5 + 2.0 ** -6) + (x.float().abs().view(xg.shape) + 1) * 2e-3
    assert bool((err <= bound).all())
    # slab vs oracle
    sl = to_np(x[1, 2]) if shape[2] > 1 else to_np(x[1, :4, 0])
    rc, rs, rm = ref.pack_lastdim(sl, g, bits)
    gc = to_np(code[1, 2]) if shape[2] > 1 else to_np(code[1, :4, 0])
    np.testing.assert_array_equal(gc, rc)


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("g", [32, 64])
def test_torch_gpu_checker_is_pinned_to_the_oracle(bits, g):
    """oracle/torch_ref.pack_lastdim (the reference's ATen chain on this GPU, used as the full-size checker in
    tests/test_reference_cases_gpu.py) == the C oracle (pinned to the reference's Python by tests/golden/), bit for bit,
    including flat groups (0/0 -> NaN -> code 0) and the dequant chain."""
    from oracle import ref, torch_ref
    rng = np.random.default_rng(bits * 100 + g)
    x = (rng.standard_normal((3, 5, 17, 256)) * rng.choice([0.01, 1.0, 30.0], size=(3, 5, 17, 1))).astype(np.float16)
    x[0, 0, 0, :g] = np.float16(0.37)                                       # a flat group
    x[1, 2, 3, g:2 * g] = 0
    c, s, m = torch_ref.pack_lastdim(torch.from_numpy(x).cuda(), g, bits)
    ec, es, em = ref.pack_lastdim(x, g, bits)
    np.testing.assert_array_equal(to_np(c), ec)
    np.testing.assert_array_equal(to_np(s).view(np.uint16), es.view(np.uint16))
    np.testing.assert_array_equal(to_np(m).view(np.uint16), em.view(np.uint16))
    d = torch_ref.unpack_dequant_lastdim(c, s, m, g, bits)
    np.testing.assert_array_equal(to_np(d).view(np.uint16), ref.unpack_dequant_lastdim(ec, es, em, g, bits).view(np.uint16))
