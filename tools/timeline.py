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
for kname, kk, cols in (("qk", 0, (0, 1, 3)), ("sv", 1, (0, 1, 2, 3))):
    a = buf[kk]
    act = a[:, 0] > 0
    rel = (a[act][:, :4].astype(np.float64) - float(t0)) / 1e3
    names = {0: "entry", 1: "first q/stats issued", 2: "blocks done", 3: "exit"}
    print(f"{kname}: {act.sum()} warps")
    for c in cols:
        x = rel[:, c]
        print(f"   {names[c]:22s} min {x.min():7.2f}  p10 {np.percentile(x, 10):7.2f}  median {np.median(x):7.2f}  p90 {np.percentile(x, 90):7.2f}  max {x.max():7.2f} us")
    d = rel[:, cols[-1]] - rel[:, 0]
    print(f"   warp lifetime          min {d.min():7.2f}  median {np.median(d):7.2f}  max {d.max():7.2f} us")
