"""Where does the multi-exit CTC kernel's time go?  (a) alpha recursion alone (grad = NULL: no beta warp, no occupancy pass), (b) the full kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "early-exit-transformer_b200")]
import torch
import eec
from eec import ops
from oracle import conformer_oracle as O
dev = torch.device("cuda")
B, T = 64, 374
lp = torch.log_softmax(torch.randn(6, B, T, 256, device=dev), -1)
tg, tl = O.synthetic_targets(B)
tg, tl = tg.to(dev), tl.to(dev)
nll = torch.empty(6, B, device=dev); loss = torch.zeros(6, device=dev); grad = torch.empty_like(lp)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=5):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
print("alpha recursion only (grad = NULL)  %.1f us" % t(lambda: ops.ctc_fwd_bwd(lp, tg, tl, nll, loss, None)))
print("full kernel + dense gradient init   %.1f us" % t(lambda: ops.ctc_fwd_bwd(lp, tg, tl, nll, loss, grad)))
for E in (1, 3):
    print("E = %d exits: full                   %.1f us" % (E, t(lambda: ops.ctc_fwd_bwd(lp[:E].contiguous(), tg, tl, nll[:E], loss[:E], grad[:E]))))
