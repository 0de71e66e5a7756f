"""A/B timing of the decode attention (kivi_decode_attention_f16) for ONE build of the library (KIVI_B200_LIB) at several
layer shapes in one process: 8 back-to-back calls in a CUDA graph (median / best of 12 replays) and single eager calls with
the L2 flushed.  Used by the tools/jobs scripts to compare tuning builds on one box, interleaved.

    KIVI_B200_LIB=$PWD/tools/variants/libkivi_X.so python tools/ab_fused.py [cfg2 cfg3 cfg4 b128]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.microbench import timeit, timeit_graph  # noqa: E402

SHAPES = {  # B, H, Hkv, T, bits, g, R
    "cfg2": (32, 32, 32, 4096, 2, 32, 128),
    "cfg3": (64, 32, 8, 8192, 2, 32, 128),
    "cfg4": (16, 32, 8, 32832, 4, 64, 64),
    "b128": (128, 32, 32, 4096, 2, 32, 128),
    "k4mha": (32, 32, 32, 4096, 4, 32, 128),      # Llama-2-7B K4V4 g32 (MHA)
    "k4g128": (32, 32, 8, 8192, 4, 128, 128),     # GQA, K4V4 g128
    "k4gqa2": (32, 32, 16, 8192, 4, 64, 64),      # ratio 2
}


def one(name, flush, gen):
    from kivi_b200.cache import KiviCache
    B, H, Hkv, T, bits, g, R = SHAPES[name]
    D, dev = 128, "cuda"
    cache = KiviCache(1, B, H, Hkv, D, bits, bits, g, R, max_tokens=T + 256, overlap_prologue=True)
    nfill = T - 1 - R // 2
    kk = torch.randn((B, Hkv, nfill, D), generator=gen, device=dev, dtype=torch.float16)
    vv = torch.randn((B, Hkv, nfill, D), generator=gen, device=dev, dtype=torch.float16)
    cache.prefill(0, kk, vv)
    del kk, vv
    qd = torch.randn((B, H, D), generator=gen, device=dev, dtype=torch.float16)
    kn = torch.randn((B, Hkv, D), generator=gen, device=dev, dtype=torch.float16)
    vn = torch.randn((B, Hkv, D), generator=gen, device=dev, dtype=torch.float16)
    out = torch.empty_like(qd)
    per_tok = D * (bits / 8 + 4 / g)
    nbytes = B * Hkv * (cache.tk * per_tok + cache.tv * per_tok + (cache.r + cache.L) * D * 2) + 2 * B * H * D * 2
    fn = lambda: cache.decode_attention(0, qd, kn, vn, out=out)  # noqa: E731
    ms, best = timeit(fn, flush=flush, iters=30)
    gms, gbest = timeit_graph(fn, reps=8, iters=12)
    r = {"shape": name, "eager_ms": round(ms, 5), "eager_best_ms": round(best, 5), "graph_ms": round(gms, 5),
         "graph_best_ms": round(gbest, 5), "graph_GBps": round(nbytes / gms / 1e6, 1)}
    del cache
    torch.cuda.empty_cache()
    return r


def main():
    names = sys.argv[1:] or ["cfg2"]
    gen = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tag = os.path.basename(os.environ.get("KIVI_B200_LIB", "default")).replace("libkivi_", "").replace(".so", "")
    for n in names:
        print(tag, json.dumps(one(n, flush, gen)), flush=True)


if __name__ == "__main__":
    main()
