"""MMA pacing probe (perf triage): time per tcgen05.mma as a function of K (MMAs per accumulator) for the v3 GEMM.
Run with EEC_LIB=.../libeec_tl.so EEC_GEMM_TL=1 [EEC_GEMM_DEBUG=38 -> no operand TMA, no staging, no activation]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
import torch
import eec
from eec import ops
dev = torch.device("cuda")
M = 23936
for K in (64, 128, 256, 512, 1024, 2048):
    for bk in (True, False):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(2048, K, device=dev) if bk else torch.randn(K, 2048, device=dev)).to(torch.bfloat16)
        o = torch.empty(M, 2048, device=dev, dtype=torch.bfloat16)
        print(f"--- K={K} B {'K-major' if bk else 'MN-major'}: {K // 16} MMAs per tile", file=sys.stderr, flush=True)
        for _ in range(2):
            ops.gemm(a, w, o, M, 2048, K, b_kmajor=bk, ldb=(K if bk else 2048))
        torch.cuda.synchronize()
