"""Per-warp timeline of one decode-attention call (cfg 2 layer shape) from a -DKIVI_TIMELINE=1 build:

    python tools/build_variants.py tl "-DKIVI_TIMELINE=1" && KIVI_B200_LIB=$PWD/tools/variants/libkivi_tl.so python tools/timeline.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kivi_b200 import _lib  # noqa: E402
from kivi_b200.cache import KiviCache  # noqa: E402

B, H, Hkv, n = 32, 32, 32, 4032
gen = torch.Generator(device="cuda").manual_seed(0)
# launched as inside a decoder layer (the q.K^T prologue may overlap the previous kernel: results are not checked here)
cache = KiviCache(1, B, H, Hkv, 128, 2, 2, 32, 128, max_tokens=n + 256,
                  overlap_prologue=True)
k = torch.randn((B, Hkv, n, 128), generator=gen, device="cuda", dtype=torch.float16)
v = torch.randn((B, Hkv, n, 128), generator=gen, device="cuda", dtype=torch.float16)
cache.prefill(0, k, v)
del k, v
q = torch.randn((B, H, 128), generator=gen, device="cuda", dtype=torch.float16)
kn = torch.randn((B, Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
vn = torch.randn((B, Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
out = torch.empty_like(q)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    cache.decode_attention(0, q, kn, vn, out=out)
torch.cuda.synchronize()
buf = np.zeros((2, 4096, 8), dtype=np.uint64)
fn = _lib.lib().kivi_debug_timeline
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
os.makedirs('gpurun_out', exist_ok=True)
np.save(os.environ.get('KIVI_TL_OUT', 'gpurun_out/timeline.npy'), buf)
t0 = buf[0][:, 0][buf[0][:, 0] > 0].min()
for kname, kk, cols in (("qk", 0, (0, 1, 2, 3)), ("sv", 1, (0, 1, 2, 3))):
    a = buf[kk]
    act = a[:, 0] > 0
    rel = (a[act][:, :4].astype(np.float64) - float(t0)) / 1e3
    names = {0: "entry", 1: "first q/stats issued", 2: "blocks done", 3: "exit"}   # qk: exit - blocks done = the ticketed cache updates
    print(f"{kname}: {act.sum()} warps")
    for c in cols:
        x = rel[:, c]
        print(f"   {names[c]:22s} min {x.min():7.2f}  p10 {np.percentile(x, 10):7.2f}  median {np.median(x):7.2f}  p90 {np.percentile(x, 90):7.2f}  max {x.max():7.2f} us")
    d = rel[:, cols[-1]] - rel[:, 0]
    print(f"   warp lifetime          min {d.min():7.2f}  median {np.median(d):7.2f}  max {d.max():7.2f} us")

# ---- where the spread of the exit times sits: between CTAs (SM position) or between the warps of a CTA
for kname, kk in (("qk", 0), ("sv", 1)):
    a = buf[kk]
    n = int((a[:, 0] > 0).sum()) // 16 * 16
    ex = ((a[:n, 3].astype(np.float64) - float(t0)) / 1e3).reshape(-1, 16)
    print(f"{kname}: exit std {ex.std():.2f} us = between CTA means {ex.mean(1).std():.2f} / within a CTA {ex.std(1).mean():.2f}; "
          f"latest warp of a CTA: median {np.median(ex.max(1)):.2f}, max {ex.max():.2f}; CTA means max {ex.mean(1).max():.2f} us")

# ---- cost regression: warp lifetime against the composition of its range (the split the library itself computes)
def _composition(kernel, n_units, n_b, n_w, w_cap):
    fn = _lib.bind("kivi_debug_range_split", ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p])
    per_unit = n_b + n_w + 1
    lo = np.zeros((w_cap + 2, 2), np.int32)
    owner = np.zeros(n_units * per_unit, np.int32)
    W = fn(n_units, n_b, n_w, w_cap, kernel, lo.ctypes.data, owner.ctypes.data)
    j = np.tile(np.arange(per_unit), n_units)
    unit = np.repeat(np.arange(n_units), per_unit)
    nb = np.bincount(owner, weights=(j < n_b), minlength=W)
    nw = np.bincount(owner, weights=((j >= n_b) & (j < n_b + n_w)), minlength=W)
    nn = np.bincount(owner, weights=(j == per_unit - 1), minlength=W)
    visits = np.array([len(np.unique(unit[owner == w])) for w in range(W)], float)
    return W, np.c_[nb, nw, nn, visits]


R = cache.residual_length
n_units = B * Hkv
w_cap = int(act.sum())
for kname, kk, (n_b, n_w), c0 in (("qk", 0, (cache.tk // 128, -(-cache.r // 16)), 0), ("sv", 1, (-(-cache.tv // 128), -(-cache.L // 16)), 1)):
    W, X = _composition(kk, n_units, n_b, n_w, w_cap)
    a = buf[kk]
    life = (a[:W, 3].astype(np.float64) - a[:W, c0].astype(np.float64)) / 1e3
    # items sum to ~const per warp: regress on (window items, new tokens, visits) with blocks absorbed by "per item"
    A = np.c_[X.sum(1) - X[:, 3], X[:, 1], X[:, 2], X[:, 3]]          # [items, of which window, of which new, visits]
    coef, *_ = np.linalg.lstsq(A, life, rcond=None)
    blk = coef[0]
    print(f"{kname}: {W} ranges, items/range {A[:, 0].min():.0f}..{A[:, 0].max():.0f}; lifetime = {blk:.3f} us/block, window item "
          f"{(blk + coef[1]) / blk:.2f} blocks, new token {(blk + coef[2]) / blk:.2f}, visit {coef[3] / blk:.2f}; "
          f"residual std {np.std(life - A @ coef):.2f} us, lifetime std {life.std():.2f} us, max - median {life.max() - np.median(life):.2f} us")
