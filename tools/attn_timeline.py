"""Per-key-block timeline of the attention forward kernel (CTA 0), from the in-kernel clock64 stamps of the `make attn_timeline` build:
    make -C early-exit-transformer_b200 attn_timeline && EEC_LIB=early-exit-transformer_b200/eec/libeec_tl.so python tools/attn_timeline.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "early-exit-transformer_b200")]
import torch
import eec
from eec import ops
B, T, H = 64, 374, 8
qkv = (torch.randn(B * T, 768, device="cuda")).to(torch.bfloat16)
kl = torch.full((B,), T, dtype=torch.int32, device="cuda")
ctx = torch.empty(B * T, 256, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
for _ in range(3):
    ops.attn_fwd(qkv, kl, ctx, lse, B, T, H)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 64))()
lib = eec.load()
lib.eec_debug_attn_timeline.argtypes = [C.c_void_p]
assert lib.eec_debug_attn_timeline(buf) == 0
tl = [[buf[e * 64 + i] for i in range(64)] for e in range(8)]
names = ["mma:P seen", "mma:PV issued", "sm:wait S", "sm:S ready", "sm:S in regs", "sm:max done", "sm:PV(j-1) done", "sm:P written"]
t0 = min(v for row in tl for v in row[:15] if v > 0)
print("block  " + "  ".join(f"{n:>15s}" for n in names))
for jb in range(15):
    print(f"{jb:5d}  " + "  ".join(f"{(tl[e][jb] - t0) if tl[e][jb] > 0 else -1:15d}" for e in range(8)))
print("per-block period (P written -> P written):", [tl[7][j + 1] - tl[7][j] for j in range(14)])
