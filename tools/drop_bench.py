"""Micro-benchmark of the dropout pieces at BASELINE shapes (CUDA events, L2 flushed between launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "early-exit-transformer_b200")):
    sys.path.insert(0, p)
import torch
import eec
from eec import ops
from eec.lib import ACT_SILU, ACT_DSILU

dev = "cuda"
M, F, D, B, T, H = 23936, 2048, 256, 64, 374, 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


st = torch.tensor([1, 0], dtype=torch.int64, device=dev)
d0 = ops.Drop(st, 0.1, 5)
print("bits FFN  [M,2048] W16 : %.1f us" % timeit(lambda: d0.with_bits(M, F, F, 16)))
print("bits ln3  [M,256]  W32 : %.1f us" % timeit(lambda: d0.with_bits(M, D, D, 32)))
print("bits attn [BHT,T]  W32 : %.1f us" % timeit(lambda: d0.with_bits(B * H * T, T, 376, 32)))
a = torch.randn(M, D, device=dev).bfloat16(); w = (torch.randn(F, D, device=dev) * 0.05).bfloat16(); bias = torch.zeros(F, device=dev)
out = torch.empty(M, F, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(out)
dF = d0.with_bits(M, F, F, 16)
print("FFN-up SiLU+pre        : %.1f us" % timeit(lambda: ops.gemm(a, w, out, M, F, D, bias=bias, act=ACT_SILU, preact=pre)))
print("FFN-up SiLU+pre + drop : %.1f us" % timeit(lambda: ops.gemm(a, w, out, M, F, D, bias=bias, act=ACT_SILU, preact=pre, drop=dF)))
dy = torch.randn(M, D, device=dev).bfloat16(); w2 = (torch.randn(D, F, device=dev) * 0.05).bfloat16(); dh = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
kw = dict(a_kmajor=True, b_kmajor=False, lda=D, ldb=F, act=ACT_DSILU, preact=pre, alpha=0.5)
print("dSiLU dgrad            : %.1f us" % timeit(lambda: ops.gemm(dy, w2, dh, M, F, D, **kw)))
print("dSiLU dgrad + drop     : %.1f us" % timeit(lambda: ops.gemm(dy, w2, dh, M, F, D, drop=dF, **kw)))
qkv = torch.randn(M, 768, device=dev).bfloat16(); kl = torch.full((B,), T, dtype=torch.int32, device=dev)
ctx = torch.empty(M, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
dA = ops.attn_drop(d0, qkv, B, T, H)
print("attn fwd               : %.1f us" % timeit(lambda: ops.attn_fwd(qkv, kl, ctx, lse, B, T, H)))
print("attn fwd + drop        : %.1f us" % timeit(lambda: ops.attn_fwd(qkv, kl, ctx, lse, B, T, H, drop=dA)))
dctx = torch.randn(M, D, device=dev).bfloat16(); dqkv = torch.empty_like(qkv); dvec = torch.empty(B * H * T, device=dev); dq32 = torch.empty(M, D, device=dev)
print("attn bwd               : %.1f us" % timeit(lambda: ops.attn_bwd(qkv, ctx, dctx, lse, kl, dqkv, dvec, B, T, H, dq32), n=5))
print("attn bwd + drop        : %.1f us" % timeit(lambda: ops.attn_bwd(qkv, ctx, dctx, lse, kl, dqkv, dvec, B, T, H, dq32, drop=dA), n=5))
res = torch.randn(M, D, device=dev); x = torch.empty(M, D, device=dev); ln = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
g = torch.ones(D, device=dev); bb = torch.zeros(D, device=dev); wl = (torch.randn(D, D, device=dev) * 0.05).bfloat16(); bl = torch.zeros(D, device=dev)
dL = d0.with_bits(M, D, D, 32)
print("ln3 K=256              : %.1f us" % timeit(lambda: ops.gemm(a, wl, x, M, D, D, bias=bl, residual=res, ln_gamma=g, ln_beta=bb, ln_out=ln)))
print("ln3 K=256 + drop       : %.1f us" % timeit(lambda: ops.gemm(a, wl, x, M, D, D, bias=bl, residual=res, ln_gamma=g, ln_beta=bb, ln_out=ln, drop=dL)))
