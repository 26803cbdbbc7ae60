"""Per-kernel micro-benchmarks at BASELINE shapes (CUDA events, L2 flushed between launches).
Usage: python tools/kbench.py [attn] [gemm] [ffn] [conv] [ctc] ...   -> prints one line per kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
import torch
import eec
from eec import ops

dev = torch.device("cuda")
B, T, H, D, F = 64, 374, 8, 256, 2048
N = B * T
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


PROFILE = os.environ.get("KBENCH_PROFILE") == "1"   # under `ncu --profile-from-start off`: exactly ONE profiled call per kernel


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    if PROFILE:
        torch.cuda.synchronize()
        flush.zero_()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return float("nan")
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def report(name, ms, flops=None, bytes_=None):
    s = f"{name:58s} {ms*1e3:9.1f} us"
    if flops: s += f"  {flops/ms/1e9:8.1f} TFLOP/s"
    if bytes_: s += f"  {bytes_/ms/1e6:8.1f} GB/s"
    print(s, flush=True)


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


which = set(sys.argv[1:]) or {"attn", "gemm", "conv", "ctc", "ln"}

if "attn" in which:
    qkv = bf(N, 768)
    kl = torch.full((B,), T, dtype=torch.int32, device=dev)
    ctx = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    fl = 4.0 * B * H * T * T * 32
    report("attn_fwd (EEC_ATTN_TC=%s)" % os.environ.get("EEC_ATTN_TC", "0"), timeit(lambda: ops.attn_fwd(qkv, kl, ctx, lse, B, T, H)), fl)
    dctx = bf(N, D); dqkv = torch.empty_like(qkv); dvec = torch.empty(B * H * T, device=dev); dq32 = torch.empty(N, D, device=dev)
    report("attn_bwd", timeit(lambda: ops.attn_bwd(qkv, ctx, dctx, lse, kl, dqkv, dvec, B, T, H, dq32), n=3, warm=1), 2.5 * fl)

if "gemm" in which:
    a256, a2048, a768, a512 = bf(N, 256), bf(N, 2048), bf(N, 768), bf(N, 512)
    w = {k: bf(*k, scale=0.05) for k in [(2048, 256), (256, 2048), (768, 256), (256, 256), (512, 256)]}
    bias = {n: torch.zeros(n, device=dev) for n in (256, 512, 768, 2048)}
    o2048 = torch.empty(N, 2048, device=dev, dtype=torch.bfloat16); pre2048 = torch.empty_like(o2048)
    o768 = torch.empty(N, 768, device=dev, dtype=torch.bfloat16)
    o256 = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    x32 = torch.randn(N, 256, device=dev); xo = torch.empty_like(x32); u = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    g = 1 + 0.1 * torch.randn(256, device=dev); b_ = torch.zeros(256, device=dev)
    report("gemm ffn1 256->2048 +bias+SiLU (bf16 out)", timeit(lambda: ops.gemm(a256, w[(2048, 256)], o2048, N, 2048, 256, bias=bias[2048], act=ops.ACT_SILU)), 2.0 * N * 2048 * 256)
    report("gemm ffn1 256->2048 +bias+SiLU +preact", timeit(lambda: ops.gemm(a256, w[(2048, 256)], o2048, N, 2048, 256, bias=bias[2048], act=ops.ACT_SILU, preact=pre2048)), 2.0 * N * 2048 * 256)
    report("gemm ffn2 2048->256 +res+LN tail", timeit(lambda: ops.gemm(a2048, w[(256, 2048)], xo, N, 256, 2048, bias=bias[256], alpha=0.5, residual=x32, ln_gamma=g, ln_beta=b_, ln_out=u)), 2.0 * N * 2048 * 256)
    report("gemm qkv 256->768 +bias", timeit(lambda: ops.gemm(a256, w[(768, 256)], o768, N, 768, 256, bias=bias[768])), 2.0 * N * 768 * 256)
    report("gemm out-proj 256->256 +res+LN", timeit(lambda: ops.gemm(a256, w[(256, 256)], xo, N, 256, 256, bias=bias[256], residual=x32, ln_gamma=g, ln_beta=b_, ln_out=u)), 2.0 * N * 256 * 256)
    report("gemm pw1 256->512 GLU", timeit(lambda: ops.gemm(a256, w[(512, 256)], o256, N, 512, 256, bias=bias[512], act=ops.ACT_GLU)), 2.0 * N * 512 * 256)
    dW = torch.zeros(2048, 256, device=dev)
    report("wgrad [N,2048]^T[N,256] split-K accumulate", timeit(lambda: ops.gemm(a2048, a256, dW, 2048, 256, N, a_kmajor=False, b_kmajor=False, lda=2048, ldb=256, accumulate=True)), 2.0 * N * 2048 * 256)
    dh = torch.empty(N, 2048, device=dev, dtype=torch.bfloat16)
    report("dgrad [N,256]x[256,2048] +dSiLU", timeit(lambda: ops.gemm(a256, w[(256, 2048)], dh, N, 2048, 256, a_kmajor=True, b_kmajor=False, lda=256, ldb=2048, act=ops.ACT_DSILU, preact=pre2048, alpha=0.5)), 2.0 * N * 2048 * 256)
    du = torch.empty(N, 256, device=dev)
    report("dgrad [N,2048]x[2048,256] fp32 out", timeit(lambda: ops.gemm(a2048, w[(2048, 256)], du, N, 256, 2048, a_kmajor=True, b_kmajor=False, lda=2048, ldb=256)), 2.0 * N * 2048 * 256)

if "ffn" in which:
    uu = bf(N, 256); w1 = bf(2048, 256, scale=0.05); w2 = bf(256, 2048, scale=0.02)
    b1 = torch.randn(2048, device=dev) * 0.1; b2 = torch.randn(256, device=dev) * 0.1
    xin = torch.randn(N, 256, device=dev); xo = torch.empty_like(xin); lo = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    g = 1 + 0.1 * torch.randn(256, device=dev); b_ = 0.1 * torch.randn(256, device=dev)
    hp = torch.empty(N, 2048, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * N * 2048 * 256
    report("ffn fused fwd (inference: no pre-activation store)", timeit(lambda: ops.ffn_fwd(uu, w1, b1, w2, b2, xin, 0.5, g, b_, xo, lo)), fl)
    report("ffn fused fwd (training: + bf16 pre-activation store)", timeit(lambda: ops.ffn_fwd(uu, w1, b1, w2, b2, xin, 0.5, g, b_, xo, lo, hpre=hp)), fl)
    # reference: plain torch in fp32 on the same bf16-rounded operands
    h = uu.float() @ w1.float().t() + b1
    a = (h * torch.sigmoid(h)).to(torch.bfloat16).float()
    xr = xin + 0.5 * (a @ w2.float().t() + b2)
    lr = torch.nn.functional.layer_norm(xr, (256,), g, b_, 1e-5)
    print("   max|x_out - ref| = %.3e (ref max %.3e)   max|ln - ref| = %.3e   max|hpre - ref| = %.3e" % (
        (xo - xr).abs().max().item(), xr.abs().max().item(), (lo.float() - lr).abs().max().item(), (hp.float() - h).abs().max().item()), flush=True)

if "conv" in which:
    gg = bf(B, T, 256); wdw = torch.randn(256, 31, device=dev) * 0.1; z = torch.zeros(256, device=dev); one = torch.ones(256, device=dev)
    out = torch.empty(B, T, 256, device=dev, dtype=torch.bfloat16)
    report("dwconv+BN+SiLU eval (bf16)", timeit(lambda: ops.dwconv_bn_silu_eval(gg, wdw, z, one, z, z, one, out, B, T, 31)), None, N * 256 * 4)
    c = torch.empty(N, 256, device=dev); sums = torch.zeros(512, dtype=torch.float64, device=dev)
    report("dwconv + stats (train pass A)", timeit(lambda: ops.dwconv_stats(gg, wdw, z, c, sums, B, T, 31)), None, N * 256 * 6)

    dc = torch.randn(N, 256, device=dev); dg = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(256, 31, device=dev); dbias = torch.zeros(256, device=dev)
    report("dwconv bwd (data + wgrad, 2 launches)", timeit(lambda: ops.dwconv_bwd(dc, gg, wdw, dg, dw, dbias, B, T, 31)), None, N * 256 * (4 + 2 + 4 + 2))
    s_ = torch.empty(N, 256, device=dev, dtype=torch.bfloat16); sm = torch.empty(256, device=dev); sr = torch.empty(256, device=dev)
    rm = torch.zeros(256, device=dev); rv = torch.ones(256, device=dev); nbt = torch.zeros((), dtype=torch.int64, device=dev)
    sums.zero_(); ops.dwconv_stats(gg, wdw, z, c, sums, B, T, 31)
    report("bn_silu_train", timeit(lambda: ops.bn_silu_train(c, sums, one, z, rm, rv, nbt, 0.1, sm, sr, s_)), None, N * 256 * 6)
    sums2 = torch.zeros(512, dtype=torch.float64, device=dev); dgam = torch.zeros(256, device=dev); dbet = torch.zeros(256, device=dev)
    def bnb():
        sums2.zero_(); ops.bn_silu_bwd(dg, c, sm, sr, one, z, sums2, dc, dgam, dbet)
    report("bn_silu_bwd (stats + apply, 2 launches + memset)", timeit(bnb), None, N * 256 * (6 + 6 + 4))
    zz = bf(N, 512); dz = torch.empty_like(zz)
    report("glu_bwd", timeit(lambda: ops.glu_bwd(zz, dg, dz)), None, N * 256 * (4 + 2 + 4))

if "ln" in which:
    x = torch.randn(N, 256, device=dev); g = torch.ones(256, device=dev); b_ = torch.zeros(256, device=dev)
    o = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    report("layernorm fwd fp32->bf16", timeit(lambda: ops.layernorm_fwd(x, g, b_, o)), None, N * 256 * 6)

if "lnbwd" in which:
    x = torch.randn(N, 256, device=dev); dy = torch.randn(N, 256, device=dev)
    g = torch.ones(256, device=dev); b_ = torch.zeros(256, device=dev)
    o = torch.empty(N, 256, device=dev, dtype=torch.bfloat16); mean = torch.empty(N, device=dev); rstd = torch.empty(N, device=dev)
    ops.layernorm_fwd(x, g, b_, o, mean, rstd)
    dx = torch.zeros(N, 256, device=dev); dcopy = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    dg_, db_, cs_ = torch.zeros(256, device=dev), torch.zeros(256, device=dev), torch.zeros(256, device=dev)
    report("layernorm bwd (+bf16 copy, dgamma/dbeta, colsum)", timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dx, True, dg_, db_, dcopy, cs_, 1.0)), None, N * 256 * 18 + N * 8)
    dyh = dy.to(torch.bfloat16)
    report("layernorm bwd, bf16 upstream gradient (the in-step form)", timeit(lambda: ops.layernorm_bwd(dyh, x, mean, rstd, g, dx, True, dg_, db_, dcopy, cs_, 1.0)), None, N * 256 * 16 + N * 8)

if "beam" in which:
    lp6 = torch.log_softmax(torch.randn(6, B, T, 256, device=dev), -1)
    dec = eec.cuda_ctc_decoder([str(i) for i in range(256)], nbest=1, beam_size=10, blank_skip_threshold=0.95)
    report("ctc prefix beam search, 6 exits x 64 utts, beam 10", timeit(lambda: dec.search(lp6), n=5), None, lp6.numel() * 4)

if "dec" in which:
    Ld = 82
    Nd = B * Ld
    qkv_d = bf(Nd, 768); kv_e = bf(N, 3072); q_d = bf(Nd, 256)
    toks = torch.randint(3, 120, (B, Ld), device=dev); toks[:, 60:] = 126
    bits = ops.key_bits_from_tokens(toks, 126)
    ctx_d = torch.empty(Nd, 256, device=dev, dtype=torch.bfloat16); lse_d = torch.empty(B, H, Ld, device=dev)
    report("decoder causal self-attention fwd (L = 82)", timeit(lambda: ops.attn_general_fwd(qkv_d[:, :256], qkv_d[:, 256:512], qkv_d[:, 512:], ctx_d, lse_d, B, Ld, Ld, H, key_bits=bits, causal=True)), 4.0 * B * H * Ld * Ld * 32)
    report("decoder cross-attention fwd (82 x 374)", timeit(lambda: ops.attn_general_fwd(q_d, kv_e[:, 512:768], kv_e[:, 768:1024], ctx_d, lse_d, B, Ld, T, H)), 4.0 * B * H * Ld * T * 32)
    dctx_d = bf(Nd, 256); dq_d = torch.empty(Nd, 256, device=dev, dtype=torch.bfloat16); dkv_e = torch.empty(N, 512, device=dev, dtype=torch.bfloat16)
    report("decoder cross-attention bwd (82 x 374)", timeit(lambda: ops.attn_general_bwd(q_d, kv_e[:, 512:768], kv_e[:, 768:1024], ctx_d, dctx_d, lse_d, dq_d, dkv_e[:, :256], dkv_e[:, 256:], B, Ld, T, H), n=3, warm=1), 10.0 * B * H * Ld * T * 32)

if "ctc" in which:
    from oracle import conformer_oracle as O
    lp = torch.log_softmax(torch.randn(6, B, T, 256, device=dev), -1)
    tg, tl = O.synthetic_targets(B)
    tg, tl = tg.to(dev), tl.to(dev)
    nll = torch.empty(6, B, device=dev); loss = torch.zeros(6, device=dev); grad = torch.empty_like(lp)
    report("ctc fwd+bwd 6 exits x 64 utts", timeit(lambda: ops.ctc_fwd_bwd(lp, tg, tl, nll, loss, grad), n=5), None, lp.numel() * 8)
    lp1 = lp[:1].contiguous(); nll1 = torch.empty(1, B, device=dev); loss1 = torch.zeros(1, device=dev); grad1 = torch.empty_like(lp1)
    report("ctc fwd+bwd 1 exit x 64 utts (serial-chain share)", timeit(lambda: ops.ctc_fwd_bwd(lp1, tg, tl, nll1, loss1, grad1), n=5), None, lp1.numel() * 8)
