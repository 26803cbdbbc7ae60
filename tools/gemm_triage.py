"""Perf triage of the K = 256 activation GEMMs (FFN up-projection forward, dSiLU data gradient) at BASELINE shapes.
  python tools/gemm_triage.py [silu|silu_pre|dsilu|qkv|all]      CUDA-event time (L2 flushed), honours EEC_GEMM_DEBUG / EEC_GEMM_TL / EEC_LIB
  KBENCH_PROFILE=1 ... under `ncu --profile-from-start off`: exactly one profiled launch per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
import torch
import eec
from eec import ops

dev = torch.device("cuda")
N = 64 * 374
PROFILE = os.environ.get("KBENCH_PROFILE") == "1"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if PROFILE:
        flush.zero_(); torch.cuda.synchronize()
        torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
        return float("nan")
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def bf(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(torch.bfloat16)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
a256 = bf(N, 256); w1 = bf(2048, 256, scale=0.05); w2 = bf(256, 2048, scale=0.05); w3 = bf(768, 256, scale=0.05)
b1 = torch.randn(2048, device=dev) * 0.1; b3 = torch.zeros(768, device=dev)
o = torch.empty(N, 2048, device=dev, dtype=torch.bfloat16); pre = bf(N, 2048); o768 = torch.empty(N, 768, device=dev, dtype=torch.bfloat16)
fl = 2.0 * N * 2048 * 256
tag = "dbg=%s" % os.environ.get("EEC_GEMM_DEBUG", "0")
cases = {
    "silu": lambda: ops.gemm(a256, w1, o, N, 2048, 256, bias=b1, act=ops.ACT_SILU),
    "silu_pre": lambda: ops.gemm(a256, w1, o, N, 2048, 256, bias=b1, act=ops.ACT_SILU, preact=pre),
    "dsilu": lambda: ops.gemm(a256, w2, o, N, 2048, 256, a_kmajor=True, b_kmajor=False, lda=256, ldb=2048, act=ops.ACT_DSILU, preact=pre, alpha=0.5),
    "qkv": lambda: ops.gemm(a256, w3, o768, N, 768, 256, bias=b3),
}
for name, fn in cases.items():
    if which not in ("all", name):
        continue
    ms = timeit(fn)
    f = fl if name != "qkv" else 2.0 * N * 768 * 256
    print(f"{name:10s} {tag:10s} {ms * 1e3:8.1f} us  {f / ms / 1e9:7.1f} TFLOP/s", flush=True)
