"""CTC timing split: alpha chain only (no gradient) vs full forward-backward, 1 and 6 exits (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
import torch
import eec
from eec import ops
from oracle import conformer_oracle as O
dev = torch.device("cuda")
B, T = 64, 374
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
tg, tl = O.synthetic_targets(B)
print("target lengths: min %d max %d mean %.1f" % (int(tl.min()), int(tl.max()), float(tl.float().mean())))
tg, tl = tg.to(dev), tl.to(dev)
for E in (1, 6):
    lp = torch.log_softmax(torch.randn(E, B, T, 256, device=dev), -1)
    nll = torch.empty(E, B, device=dev); loss = torch.zeros(E, device=dev); grad = torch.empty_like(lp)
    print(f"E={E}: loss only (alpha chain) {timeit(lambda: ops.ctc_fwd_bwd(lp, tg, tl, nll, loss, None)):7.1f} us   fwd+bwd {timeit(lambda: ops.ctc_fwd_bwd(lp, tg, tl, nll, loss, grad)):7.1f} us")
