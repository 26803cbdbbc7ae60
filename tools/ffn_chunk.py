"""Does the [rows, 2048] FFN activation survive in L2 between the up- and the down-projection when the feed-forward module is run in
row chunks?  Times FFN-up (+SiLU) -> FFN-down (+residual + LayerNorm tail) over N = 23 936 rows in 1 / 2 / 3 / 4 / 6 row chunks
(CUDA events, L2 flushed before each measurement); `train` also stores the pre-activation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
import torch
import eec
from eec import ops
dev = torch.device("cuda")
N, D, F = 64 * 374, 256, 2048
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bf(*s, scale=1.0): return (torch.randn(*s, device=dev) * scale).to(torch.bfloat16)
u = bf(N, D); w1 = bf(F, D, scale=0.05); w2 = bf(D, F, scale=0.02)
b1 = torch.randn(F, device=dev) * 0.1; b2 = torch.randn(D, device=dev) * 0.1
xin = torch.randn(N, D, device=dev); xo = torch.empty_like(xin); lo = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
g = 1 + 0.1 * torch.randn(D, device=dev); be = 0.1 * torch.randn(D, device=dev)
act = torch.empty(N, F, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(act)
def run(chunks, train, shared):
    step = (N + chunks - 1) // chunks
    step = (step + 255) // 256 * 256
    for r0 in range(0, N, step):
        r1 = min(N, r0 + step); m = r1 - r0
        a = act[:m] if shared else act[r0:r1]          # shared: every chunk reuses the same activation rows (inference: nothing to keep)
        ops.gemm(u[r0:r1], w1, a, m, F, D, bias=b1, act=ops.ACT_SILU, preact=(pre[r0:r1] if train else None))
        ops.gemm(a, w2, xo[r0:r1], m, D, F, bias=b2, alpha=0.5, residual=xin[r0:r1], ln_gamma=g, ln_beta=be, ln_out=lo[r0:r1])
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()                      # one graph replay: the host cost of the ctypes launches is not part of the answer
    with torch.cuda.graph(gr):
        fn()
    fn = gr.replay
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
ref = None
for train in (False, True):
    for chunks in (1, 2, 3, 4, 6):
        for shared in ((False, True) if not train else (False,)):
            t = timeit(lambda: run(chunks, train, shared))
            run(chunks, train, shared); torch.cuda.synchronize()
            if ref is None: ref = xo.clone()
            err = float((xo - ref).abs().max())
            print(f"{'train' if train else 'infer'} chunks {chunks} {'shared act buffer' if shared else 'full act tensor  '}: {t:7.1f} us per FFN   max|x - x(1 chunk)| = {err:.1e}", flush=True)
