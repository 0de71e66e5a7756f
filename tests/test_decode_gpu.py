"""Parity of the pre-allocated cache + fused decode-attention kernel (kivi_cache.cu, kivi_decode.cu)
with the restated attention hook of the reference (oracle/ref.py: models/llama_kivi.py:314-455).

Cache contents (codes, scale, mn, fp16 windows) are compared BIT-EXACTLY with the oracle's 9-tuple.
The attention output passes through the reference's fp16 rounding points (fp16 logits -> fp16 scale
-> fp32 softmax -> fp16 probs -> fp16 partial outputs), where a 1-ulp flip of an fp16 logit (ulp up
to 2^-7 at |s| ~ 8) legitimately moves a probability by ~1%; so every stage is checked against the
oracle applied to the kernel's OWN previous-stage values (rtol 1e-3 + fp32 accumulation floor), and
the end-to-end output against the full oracle chain with the looser, stated E2E tolerance."""
import numpy as np
import pytest
import torch

from oracle import ref
from tests._util import assert_gemv_close, l1_mass_ref_layout, to_np

pytestmark = pytest.mark.gpu

E2E_RTOL, E2E_ATOL_FRAC = 2e-2, 5e-3       # end-to-end |err| <= 2e-2*|ref| + 5e-3*max|ref|


def _tuple_equal(got, exp):
    assert got[8] == exp[8]
    for i in range(8):
        a, b = got[i], exp[i]
        if b is None or b.size == 0:
            assert a is None or a.numel() == 0, f"tuple[{i}] should be empty"
            continue
        a = to_np(a)
        assert a.shape == b.shape, (i, a.shape, b.shape)
        if a.dtype == np.float16:
            np.testing.assert_array_equal(a.view(np.uint16), b.view(np.uint16), err_msg=f"tuple[{i}]")
        else:
            np.testing.assert_array_equal(a, b, err_msg=f"tuple[{i}]")


def _mk_cache(B, H, Hkv, kb, vb, g, R, max_tokens=1024, n_layers=1, mode=None):
    from kivi_b200.cache import KiviCache
    return KiviCache(n_layers, B, H, Hkv, 128, kb, vb, g, R, max_tokens)


@pytest.fixture(params=["G-auto", "G-1"])
def mode(request, monkeypatch):
    """G-auto: the query heads of a KV head share the MMAs (chunks of up to 4); G-1: one head per unit."""
    if request.param == "G-1":
        monkeypatch.setenv("KIVI_GQA_G", "1")
    else:
        monkeypatch.delenv("KIVI_GQA_G", raising=False)
    return request.param


@pytest.mark.parametrize("n", [1, 5, 128, 200, 333, 640])
@pytest.mark.parametrize("kb,vb,g,R", [(2, 2, 32, 128), (4, 4, 64, 64), (2, 4, 32, 32), (4, 2, 128, 128)])
def test_prefill_matches_oracle(n, kb, vb, g, R):
    """kivi_cache_prefill_f16 == the prefill split + pack of models/llama_kivi.py:425-452, bit for bit."""
    rng = np.random.default_rng(n * 31 + kb + R)
    B, H, Hkv = 2, 4, 2
    k = rng.standard_normal((B, Hkv, n, 128)).astype(np.float16)
    v = rng.standard_normal((B, Hkv, n, 128)).astype(np.float16)
    cache = _mk_cache(B, H, Hkv, kb, vb, g, R)
    cache.prefill(0, torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda())
    exp = list(ref.prefill_cache(k, v, g, kb, vb, R))
    # the oracle packs K and V with their own bit widths
    if exp[0] is not None:
        nq = exp[0].shape[-1] * (32 // kb)
        exp[0], exp[2], exp[3] = ref.pack_lastdim(np.ascontiguousarray(k[:, :, :nq].transpose(0, 1, 3, 2)), g, kb)
    if exp[4] is not None:
        exp[4], exp[6], exp[7] = ref.pack_lastdim(np.ascontiguousarray(v[:, :, :-R]), g, vb)
    _tuple_equal(cache.export(0), tuple(exp))
    assert to_np(cache.state)[:6].tolist() == [cache.tk, cache.r, cache.tv, cache.L, cache.vhead, cache.kv_len]


def _stage_checks(cache_tuple_before, q, k_new, v_new, g, kb, vb, R, got_out, got_s, got_p, mask=None):
    """cache_tuple_before: oracle 9-tuple BEFORE the step (numpy)."""
    Kq, Kfull, Ks, Kz, Vq, Vfull, Vs, Vz, kv_len = cache_tuple_before
    B, H, _, D = q.shape
    T = kv_len + 1
    # ---- stage 1: logits (fp16 kernel outputs), then the fp16 scale
    Kf = np.concatenate([Kfull, k_new], axis=2) if Kfull is not None else k_new
    parts, l1 = [], []
    if Kq is not None:
        parts.append(ref.bmm_fA_qB_outer(g, q, Kq, Ks, Kz, kb))
        l1.append(np.broadcast_to(l1_mass_ref_layout(q, Ks, Kz, 2 ** kb - 1), parts[-1].shape))
    parts.append(ref.residual_qk(q, Kf))
    rep = H // Kf.shape[1]
    l1r = np.einsum("bhd,bhtd->bht", np.abs(q[:, :, 0].astype(np.float64)),
                    np.abs(np.repeat(Kf, rep, axis=1).astype(np.float64)))[:, :, None, :]
    l1.append(l1r)
    logits = np.concatenate(parts, -1)
    l1 = np.concatenate(l1, -1)
    exp_s = (logits.astype(np.float32) * (np.float32(1.0) / np.float32(11.313708))).astype(np.float16)
    if mask is not None:
        exp_s = (exp_s.astype(np.float32) + mask.astype(np.float32)).astype(np.float16)
        exp_s = np.maximum(exp_s, np.float16(-65504))
    # The kernel output that the 1e-3 rtol bar applies to is the UNSCALED fp16 logit (the reference
    # kernel's output); the fp16 scale that follows re-rounds it.  Accept exactly the scaled images of
    # the oracle logit and of its two fp16 neighbours (a 1-ulp flip = 2^-10 relative <= 1e-3), or the
    # fp32 accumulation floor for logits that cancel to ~0.
    def _sc(x):
        y = (x.astype(np.float32) * (np.float32(1.0) / np.float32(11.313708))).astype(np.float16)
        if mask is not None:
            y = np.maximum((y.astype(np.float32) + mask.astype(np.float32)).astype(np.float16), np.float16(-65504))
        return y
    gs = got_s[..., :T]
    ok = np.zeros(gs.shape, bool)
    for cand in (logits, np.nextafter(logits, np.float16(-np.inf)), np.nextafter(logits, np.float16(np.inf))):
        ok |= (gs == _sc(cand))
    ok |= np.abs(gs.astype(np.float64) - exp_s.astype(np.float64)) <= 1e-6 * l1 / 11.3
    assert ok.all(), f"scaled logits: {(~ok).sum()} / {ok.size} differ by more than one fp16 ulp of the kernel output"
    # ---- stage 2: softmax of the kernel's own scaled logits
    exp_p = ref.scale_softmax(np.ascontiguousarray(got_s[..., :T]), 1)
    pe = np.abs(got_p[..., :T].astype(np.float64) - exp_p.astype(np.float64))
    assert (pe <= 1e-3 * exp_p.astype(np.float64) + 1e-7).all(), f"softmax stage: max err {pe.max():.3e}"
    # ---- stage 3: p.V with the kernel's own probabilities
    p_own = np.ascontiguousarray(got_p[..., :T])
    Vf = np.concatenate([Vfull, v_new], axis=2)
    L = Vf.shape[2]
    out_r = ref.residual_pv(np.ascontiguousarray(p_own[..., -L:]), Vf)
    l1o = np.einsum("bht,bhtd->bhd", np.abs(p_own[:, :, 0, -L:].astype(np.float64)),
                    np.abs(np.repeat(Vf, rep, axis=1).astype(np.float64)))[:, :, None, :]
    if Vq is not None:
        pq = np.ascontiguousarray(p_own[..., :-L])
        out_q = ref.bmm_fA_qB_outer(g, pq, Vq, Vs, Vz, vb)
        exp_out = ref.add_f16(out_q, out_r)
        l1o = l1o + l1_mass_ref_layout(pq, Vs, Vz, 2 ** vb - 1)
        # the two fp16 partial sums may each flip by one ulp before the fp16 add
        l1o = l1o + (2.0 ** -10 / 1e-6) * (np.abs(out_q.astype(np.float64)) + np.abs(out_r.astype(np.float64)))
    else:
        exp_out = out_r
    assert_gemv_close(got_out, exp_out, l1o, "attention output (own probs)")


DECODE_CASES = [  # B, H, Hkv, kb, vb, g, R, n_prefill, steps
    (1, 2, 2, 2, 2, 32, 128, 300, 140),     # crosses a K flush (r: 44 -> 128) and V ring wrap-around
    (2, 4, 1, 2, 2, 32, 32, 70, 40),        # GQA 4 (G = 4), R = 32: several flushes
    (1, 8, 1, 4, 4, 64, 64, 130, 70),       # ratio 8 -> two chunks of 4, 4-bit g64
    (1, 2, 1, 2, 4, 32, 64, 10, 80),        # G = 2, mixed bits, starts below R (no packed part at first)
    (1, 3, 3, 4, 2, 128, 128, 0, 135),      # decode from an EMPTY cache, g = 128
]


@pytest.mark.parametrize("B,H,Hkv,kb,vb,g,R,n0,steps", DECODE_CASES)
def test_decode_steps_match_oracle(B, H, Hkv, kb, vb, g, R, n0, steps, mode):
    rng = np.random.default_rng(n0 * 13 + H + R)
    cache = _mk_cache(B, H, Hkv, kb, vb, g, R, max_tokens=512, mode=mode)
    if n0 > 0:
        k = rng.standard_normal((B, Hkv, n0, 128)).astype(np.float16)
        v = rng.standard_normal((B, Hkv, n0, 128)).astype(np.float16)
        cache.prefill(0, torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda())
        st = list(ref.prefill_cache(k, v, g, kb, vb, R))
        if st[0] is not None:
            nq = st[0].shape[-1] * (32 // kb)
            st[0], st[2], st[3] = ref.pack_lastdim(np.ascontiguousarray(k[:, :, :nq].transpose(0, 1, 3, 2)), g, kb)
        if st[4] is not None:
            st[4], st[6], st[7] = ref.pack_lastdim(np.ascontiguousarray(v[:, :, :-R]), g, vb)
        st = tuple(st)
    else:
        st = (None, None, None, None, None, np.zeros((B, Hkv, 0, 128), np.float16), None, None, 0)
    tmax = 512
    dbg_s = torch.zeros((B, H, tmax), dtype=torch.float16, device="cuda")
    dbg_p = torch.zeros_like(dbg_s)
    worst = 0.0
    for step in range(steps):
        q = (rng.standard_normal((B, H, 1, 128)) * 0.7).astype(np.float16)
        k_new = rng.standard_normal((B, Hkv, 1, 128)).astype(np.float16)
        v_new = rng.standard_normal((B, Hkv, 1, 128)).astype(np.float16)
        out = cache.decode_attention(0, torch.from_numpy(q[:, :, 0]).cuda(), torch.from_numpy(k_new[:, :, 0]).cuda(),
                                     torch.from_numpy(v_new[:, :, 0]).cuda(), dbg_logits=dbg_s, dbg_probs=dbg_p)
        cache.advance()
        torch.cuda.synchronize()
        got_out = to_np(out)[:, :, None, :]
        got_s, got_p = to_np(dbg_s)[:, :, None, :], to_np(dbg_p)[:, :, None, :]
        check = step < 3 or step % 9 == 0 or step >= steps - 3 or st[1] is None or (st[1].shape[2] >= R - 2)
        if check:
            _stage_checks(st, q, k_new, v_new, g, kb, vb, R, got_out, got_s, got_p)
        # oracle step (K and V may use different bit widths: restate with the per-tensor widths)
        exp_out, exp_p, st = _oracle_step(st, q, k_new, v_new, g, kb, vb, R)
        err = np.abs(got_out.astype(np.float64) - exp_out.astype(np.float64))
        tol = E2E_RTOL * np.abs(exp_out.astype(np.float64)) + E2E_ATOL_FRAC * np.abs(exp_out.astype(np.float64)).max()
        assert (err <= tol).all(), f"step {step}: end-to-end err {err.max():.3e}"
        worst = max(worst, float(err.max()))
        if check:
            _tuple_equal(cache.export(0), st)
    assert to_np(cache.state)[:6].tolist() == [cache.tk, cache.r, cache.tv, cache.L, cache.vhead, cache.kv_len]


def _oracle_step(st, q, k_new, v_new, g, kb, vb, R, mask=None):
    return ref.decode_step(st, q, k_new, v_new, g, kb, vb, R, mask)


def test_decode_with_mask(mode):
    """Additive mask + max with finfo.min (models/llama_kivi.py:364-372), e.g. left padding."""
    rng = np.random.default_rng(4)
    B, H, Hkv, kb, vb, g, R, n0 = 2, 2, 2, 2, 2, 32, 128, 200
    cache = _mk_cache(B, H, Hkv, kb, vb, g, R, max_tokens=512, mode=mode)
    k = rng.standard_normal((B, Hkv, n0, 128)).astype(np.float16)
    v = rng.standard_normal((B, Hkv, n0, 128)).astype(np.float16)
    cache.prefill(0, torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda())
    st = ref.prefill_cache(k, v, g, kb, vb, R)
    T = n0 + 1
    mask = np.zeros((B, 1, 1, T), np.float16)
    mask[0, :, :, :17] = np.finfo(np.float16).min                   # sequence 0 is left-padded by 17 tokens
    mask[1, :, :, :3] = np.finfo(np.float16).min
    q = rng.standard_normal((B, H, 1, 128)).astype(np.float16)
    k_new = rng.standard_normal((B, Hkv, 1, 128)).astype(np.float16)
    v_new = rng.standard_normal((B, Hkv, 1, 128)).astype(np.float16)
    dbg_s = torch.zeros((B, H, 512), dtype=torch.float16, device="cuda")
    dbg_p = torch.zeros_like(dbg_s)
    out = cache.decode_attention(0, torch.from_numpy(q[:, :, 0]).cuda(), torch.from_numpy(k_new[:, :, 0]).cuda(),
                                 torch.from_numpy(v_new[:, :, 0]).cuda(), mask=torch.from_numpy(mask).cuda(),
                                 dbg_logits=dbg_s, dbg_probs=dbg_p)
    torch.cuda.synchronize()
    got_p = to_np(dbg_p)[:, :, None, :T]
    assert (got_p[0, :, :, :17] == 0).all() and (got_p[1, :, :, :3] == 0).all()
    _stage_checks(st, q, k_new, v_new, g, kb, vb, R, to_np(out)[:, :, None, :], to_np(dbg_s)[:, :, None, :],
                  to_np(dbg_p)[:, :, None, :], mask=np.broadcast_to(mask, (B, H, 1, T)))


def test_multi_layer_shared_state():
    """All layers share one device state; decode of layer l must not disturb layer m."""
    rng = np.random.default_rng(8)
    B, H, Hkv, g, R, n0, NL = 1, 2, 2, 32, 128, 150, 3
    cache = _mk_cache(B, H, Hkv, 2, 2, g, R, max_tokens=512, n_layers=NL)
    sts = []
    for l in range(NL):
        k = rng.standard_normal((B, Hkv, n0, 128)).astype(np.float16)
        v = rng.standard_normal((B, Hkv, n0, 128)).astype(np.float16)
        cache.prefill(l, torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda())
        sts.append(ref.prefill_cache(k, v, g, 2, 2, R))
    for step in range(4):
        for l in range(NL):
            q = rng.standard_normal((B, H, 1, 128)).astype(np.float16)
            kn = rng.standard_normal((B, Hkv, 1, 128)).astype(np.float16)
            vn = rng.standard_normal((B, Hkv, 1, 128)).astype(np.float16)
            cache.decode_attention(l, torch.from_numpy(q[:, :, 0]).cuda(), torch.from_numpy(kn[:, :, 0]).cuda(),
                                   torch.from_numpy(vn[:, :, 0]).cuda())
            _, _, sts[l] = ref.decode_step(sts[l], q, kn, vn, g, 2, 2, R)
        cache.advance()
    for l in range(NL):
        _tuple_equal(cache.export(l), sts[l])


def test_full_size_consistency(mode):
    """BASELINE cfg 2 layer shape (B32, H32, T = 4096, K2V2 g32 R128): too big for the CPU oracle end to
    end, so (1) a slab of units is checked stage-by-stage against the oracle, (2) the fused kernel must
    agree with the library's own generic-layout kernels run on the exported cache for ALL units
    (two independent code paths), (3) probabilities sum to 1."""
    from kivi_b200 import matmul
    gen = torch.Generator(device="cuda").manual_seed(3)
    B, H, Hkv, g, R, n0 = 32, 32, 32, 32, 128, 4095
    cache = _mk_cache(B, H, Hkv, 2, 2, g, R, max_tokens=4352, mode=mode)
    k = torch.randn((B, Hkv, n0, 128), generator=gen, device="cuda", dtype=torch.float16)
    v = torch.randn((B, Hkv, n0, 128), generator=gen, device="cuda", dtype=torch.float16)
    cache.prefill(0, k, v)
    del k, v
    tup = cache.export(0)
    q = torch.randn((B, H, 128), generator=gen, device="cuda", dtype=torch.float16)
    kn = torch.randn((B, Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
    vn = torch.randn((B, Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
    dbg_s = torch.zeros((B, H, 4352), dtype=torch.float16, device="cuda")
    dbg_p = torch.zeros_like(dbg_s)
    out = cache.decode_attention(0, q, kn, vn, dbg_logits=dbg_s, dbg_probs=dbg_p)
    torch.cuda.synchronize()
    T = n0 + 1
    assert cache.tk == 3968 and cache.r == 127 and cache.tv == 3967 and cache.L == 128
    # (3)
    psum = dbg_p[..., :T].float().sum(-1)
    assert bool(((psum - 1).abs() < 2e-2).all())
    # (2) generic kernels on the exported (reference-layout) cache
    lq = matmul.cuda_bmm_fA_qB_outer(g, q[:, :, None, :], tup[0], tup[2], tup[3], 2)[:, :, 0]
    s_generic = (lq.float() * (1.0 / 11.313708)).half()
    diff = (dbg_s[..., :3968].float() - s_generic.float()).abs()
    assert bool((diff <= 1e-3 * s_generic.float().abs() + 2e-3).all()), float(diff.max())
    oq = matmul.cuda_bmm_fA_qB_outer(g, dbg_p[:, :, None, :3967], tup[4], tup[6], tup[7], 2)[:, :, 0]
    vfull = torch.cat([tup[5], vn[:, :, None, :]], dim=2)
    orr = torch.matmul(dbg_p[:, :, None, 3967:T].float(), vfull.float())[:, :, 0].half()
    exp = (oq + orr)
    d2 = (out.float() - exp.float()).abs()
    assert bool((d2 <= 2e-3 * exp.float().abs() + 2e-4).all()), float(d2.max())
    # (1) oracle on a slab
    sl = slice(7, 8)
    st = tuple(None if t is None else (t if isinstance(t, int) else to_np(t[sl, :2])) for t in tup[:8]) + (tup[8],)
    _stage_checks(st, to_np(q[sl, :2])[:, :, None, :], to_np(kn[sl, :2])[:, :, None, :], to_np(vn[sl, :2])[:, :, None, :],
                  g, 2, 2, R, to_np(out[sl, :2])[:, :, None, :], to_np(dbg_s[sl, :2])[:, :, None, :],
                  to_np(dbg_p[sl, :2])[:, :, None, :])


@pytest.mark.parametrize("B,H,Hkv,kb,vb,g,R,n0", [(1, 8, 2, 4, 4, 64, 64, 20000),      # cfg-4-like: few, long units, GQA 4, 4-bit
                                                  (2, 4, 4, 2, 2, 32, 128, 9000)])     # MHA, more ranges per unit than warps of a CTA
def test_long_context_many_ranges_per_unit(B, H, Hkv, kb, vb, g, R, n0):
    """Few long units: every unit is cut into many warp ranges (more than 32 statistic slots and partial records per
    unit), far beyond what a logits row in shared memory could hold.  Checked against the library's generic-layout
    kernels run on the exported cache (an independent code path) and against the softmax identities."""
    from kivi_b200 import matmul
    gen = torch.Generator(device="cuda").manual_seed(11)
    cache = _mk_cache(B, H, Hkv, kb, vb, g, R, max_tokens=n0 + 64)
    k = torch.randn((B, Hkv, n0, 128), generator=gen, device="cuda", dtype=torch.float16)
    v = torch.randn((B, Hkv, n0, 128), generator=gen, device="cuda", dtype=torch.float16)
    cache.prefill(0, k, v)
    del k, v
    tup = cache.export(0)
    q = (torch.randn((B, H, 128), generator=gen, device="cuda", dtype=torch.float32) * 0.5).half()
    kn = torch.randn((B, Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
    vn = torch.randn((B, Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
    T = n0 + 1
    dbg_s = torch.zeros((B, H, T + 8), dtype=torch.float16, device="cuda")
    dbg_p = torch.zeros_like(dbg_s)
    out = cache.decode_attention(0, q, kn, vn, dbg_logits=dbg_s, dbg_probs=dbg_p)
    out2 = cache.decode_attention(0, q, kn, vn)                     # same state (no advance): the fast path, same result
    torch.cuda.synchronize()
    assert torch.equal(out, out2), "debug and fast paths disagree"
    tk, tv, L = cache.tk, cache.tv, cache.L
    psum = dbg_p[..., :T].float().sum(-1)
    assert bool(((psum - 1).abs() < 2e-2).all())
    lq = matmul.cuda_bmm_fA_qB_outer(g, q[:, :, None, :], tup[0], tup[2], tup[3], kb)[:, :, 0]
    s_generic = (lq.float() * (1.0 / 11.313708)).half()
    diff = (dbg_s[..., :tk].float() - s_generic.float()).abs()
    assert bool((diff <= 1e-3 * s_generic.float().abs() + 2e-3).all()), float(diff.max())
    # softmax of the kernel's own logits
    p_ref = torch.softmax(dbg_s[..., :T].float(), -1)
    assert bool(((dbg_p[..., :T].float() - p_ref).abs() <= 2e-3 * p_ref + 1e-6).all())
    oq = matmul.cuda_bmm_fA_qB_outer(g, dbg_p[:, :, None, :tv], tup[4], tup[6], tup[7], vb)[:, :, 0]
    rep = H // Hkv
    vfull = torch.cat([tup[5], vn[:, :, None, :]], dim=2).repeat_interleave(rep, dim=1)
    orr = torch.matmul(dbg_p[:, :, None, tv:T].float(), vfull.float())[:, :, 0].half()
    exp = (oq + orr)
    d2 = (out.float() - exp.float()).abs()
    assert bool((d2 <= 2e-3 * exp.float().abs() + 2e-4).all()), float(d2.max())
