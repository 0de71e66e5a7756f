"""Tiny driver for ncu: builds a cfg-2-shaped cache and runs the fused decode kernel a few times."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kivi_b200.cache import KiviCache  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--H", type=int, default=32)
ap.add_argument("--Hkv", type=int, default=32)
ap.add_argument("--n", type=int, default=4032)
ap.add_argument("--bits", type=int, default=2)
ap.add_argument("--g", type=int, default=32)
ap.add_argument("--R", type=int, default=128)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--mode", default="fused")
a = ap.parse_args()
gen = torch.Generator(device="cuda").manual_seed(0)
# launched as inside a decoder layer (the q.K^T prologue may overlap the previous kernel: results are not checked here)
cache = KiviCache(1, a.B, a.H, a.Hkv, 128, a.bits, a.bits, a.g, a.R, max_tokens=a.n + 256,
                  overlap_prologue=True)
k = torch.randn((a.B, a.Hkv, a.n, 128), generator=gen, device="cuda", dtype=torch.float16)
v = torch.randn((a.B, a.Hkv, a.n, 128), generator=gen, device="cuda", dtype=torch.float16)
cache.prefill(0, k, v)
del k, v
q = torch.randn((a.B, a.H, 128), generator=gen, device="cuda", dtype=torch.float16)
kn = torch.randn((a.B, a.Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
vn = torch.randn((a.B, a.Hkv, 128), generator=gen, device="cuda", dtype=torch.float16)
out = torch.empty_like(q)
for _ in range(a.iters):
    cache.decode_attention(0, q, kn, vn, out=out)
torch.cuda.synchronize()
print("state", cache.tk, cache.r, cache.tv, cache.L)
