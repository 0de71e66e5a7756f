"""How close are the decode-step linears (M = 32 rows) to the HBM roofline with cuBLAS, and does the operand layout matter?"""
import torch
torch.manual_seed(0)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


for M in (32,):
    for (N, K) in [(12288, 4096), (4096, 4096), (22016, 4096), (4096, 11008)]:
        x = torch.randn(M, K, device=dev, dtype=torch.float16)
        W = torch.randn(N, K, device=dev, dtype=torch.float16)          # nn.Linear layout [out, in]
        Wt = W.t().contiguous()                                         # [in, out]
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        outT = torch.empty(N, M, device=dev, dtype=torch.float16)
        xT = x.t().contiguous()
        gb = N * K * 2 / 1e6
        a = t(lambda: torch.mm(x, W.t(), out=out))
        b = t(lambda: torch.mm(x, Wt, out=out))
        c = t(lambda: torch.mm(W, xT, out=outT))                        # C^T = W x^T
        d = t(lambda: torch.nn.functional.linear(x, W))
        print(f"M={M} N={N} K={K}: mm(x,W.t) {a*1e3:.1f} us {gb/a:.0f} GB/s | mm(x,Wt) {b*1e3:.1f} us {gb/b:.0f} | mm(W,xT) {c*1e3:.1f} us {gb/c:.0f} | linear {d*1e3:.1f} us {gb/d:.0f}")
