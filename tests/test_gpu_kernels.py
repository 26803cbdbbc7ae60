"""GPU unit parity: every kernel family of libeec.so (called through the C ABI via eec.ops) against
plain fp64/fp32 PyTorch restatements of the same op on the same seeded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import eec  # noqa: F401
    from eec import ops as _ops
    eec.load()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def ref_gemm(A, B, a_k, b_k):
    A, B = A.double(), B.double()
    Am = A if a_k else A.t()
    Bm = B if b_k else B.t()
    return (Am @ Bm.t()).cpu()


MAJORS = [(True, True), (True, False), (False, False), (False, True)]


@pytest.mark.parametrize("a_k,b_k", MAJORS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_plain(ops, a_k, b_k, dtype):
    M, N, K = 384 + 40, 512, 320
    A = rnd(*((M, K) if a_k else (K, M)), seed=1, dtype=dtype)
    B = rnd(*((N, K) if b_k else (K, N)), seed=2, dtype=dtype)
    C = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, C, M, N, K, a_kmajor=a_k, b_kmajor=b_k)
    ref = ref_gemm(A, B, a_k, b_k)
    assert rel(C, ref) < (2e-6 if dtype == torch.float32 else 1e-5)  # bf16 inputs are exact products, fp32 accumulate


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_epilogues(ops, dtype):
    M, N, K = 300, 512, 256
    A, B = rnd(M, K, seed=3, dtype=dtype), rnd(N, K, seed=4, scale=0.1, dtype=dtype)
    bias, res = rnd(N, seed=5), rnd(M, N, seed=6)
    tol = 3e-6 if dtype == torch.float32 else 2e-2
    base = ref_gemm(A, B, True, True) + bias.double().cpu()
    # SiLU + preact store + bf16/fp32 out
    out = torch.empty(M, N, device="cuda", dtype=dtype)
    pre = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.gemm(A, B, out, M, N, K, bias=bias, act=ops.ACT_SILU, preact=pre)
    assert rel(pre, base) < tol
    assert rel(out, base * torch.sigmoid(base)) < tol
    # dSiLU with alpha
    out32 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, out32, M, N, K, act=ops.ACT_DSILU, preact=pre, alpha=0.5)
    h = pre.double().cpu()
    s = torch.sigmoid(h)
    assert rel(out32, 0.5 * ref_gemm(A, B, True, True) * (s * (1 + h * (1 - s)))) < tol
    # residual + alpha, row-periodic residual
    ops.gemm(A, B, out32, M, N, K, bias=bias, alpha=0.5, residual=res)
    assert rel(out32, 0.5 * base + res.double().cpu()) < tol
    pe = rnd(100, N, seed=7)
    ops.gemm(A, B, out32, M, N, K, bias=bias, residual=pe, res_row_mod=100)
    idx = torch.arange(M) % 100
    assert rel(out32, base + pe.double().cpu()[idx]) < tol
    # GLU
    z = torch.empty(M, N, device="cuda", dtype=dtype)
    gout = torch.empty(M, N // 2, device="cuda", dtype=dtype)
    ops.gemm(A, B, gout, M, N, K, bias=bias, act=ops.ACT_GLU, preact=z)
    assert rel(z, base) < tol
    assert rel(gout, base[:, : N // 2] * torch.sigmoid(base[:, N // 2:])) < tol
    # accumulate (split-K wgrad shape): C += A^T B over a long K
    Kl = 4096 + 64
    Am, Bm = rnd(Kl, 256, seed=8, scale=0.1, dtype=dtype), rnd(Kl, 512, seed=9, scale=0.1, dtype=dtype)
    Cacc = torch.ones(256, 512, device="cuda")
    ops.gemm(Am, Bm, Cacc, 256, 512, Kl, a_kmajor=False, b_kmajor=False, alpha=0.5, accumulate=True)
    assert rel(Cacc, 1.0 + 0.5 * Am.double().cpu().t() @ Bm.double().cpu()) < (1e-5 if dtype == torch.float32 else 2e-5)


@pytest.mark.parametrize("M,N", [(300, 512), (128 * 5 + 7, 2048), (23936, 768), (64, 256)])
def test_gemm_weight_stationary(ops, M, N):
    """K = 256 projections with a bf16 output run on the weight-stationary kernel (csrc/gemm_ws.cu): bias, SiLU, SiLU + stored
    pre-activation (K-major weights) and the dSiLU data gradient (MN-major weights), incl. an m-tile tail and several CTAs per
    weight tile; EEC_GEMM_WS=0 routes the same descriptors to the v3 kernel."""
    K = 256
    A = rnd(M, K, seed=31, dtype=torch.bfloat16)
    W = rnd(N, K, seed=32, scale=0.1, dtype=torch.bfloat16)
    bias = rnd(N, seed=33)
    base = ref_gemm(A, W, True, True) + bias.double().cpu()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, out, M, N, K, bias=bias)
    assert rel(out, base) < 6e-3
    out.fill_(float("nan"))
    ops.gemm(A, W, out, M, N, K, bias=bias, act=ops.ACT_SILU)
    assert rel(out, base * torch.sigmoid(base)) < 8e-3
    out.fill_(float("nan"))
    pre = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, out, M, N, K, bias=bias, act=ops.ACT_SILU, preact=pre)
    assert rel(pre, base) < 6e-3
    assert rel(out, base * torch.sigmoid(base)) < 8e-3
    # dSiLU data gradient: dH = alpha * (dY Wt) o dSiLU(pre), Wt [K, N] (MN-major B)
    Wt = rnd(K, N, seed=34, scale=0.1, dtype=torch.bfloat16)
    out.fill_(float("nan"))
    ops.gemm(A, Wt, out, M, N, K, a_kmajor=True, b_kmajor=False, lda=K, ldb=N)          # plain data gradient with a bf16 output
    assert rel(out, ref_gemm(A, Wt, True, False)) < 6e-3
    # long-K data gradient with a bf16 output (the CTA-pair streaming kernel, csrc/gemm_pair.cu): dU[M, 256] = dH[M, N] Wl[N, 256]
    if N >= 512:
        Wl = rnd(N, 256, seed=35, scale=0.05, dtype=torch.bfloat16)
        H_ = rnd(M, N, seed=36, dtype=torch.bfloat16)
        du = torch.full((M, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm(H_, Wl, du, M, 256, N, a_kmajor=True, b_kmajor=False, lda=N, ldb=256)
        assert rel(du, ref_gemm(H_, Wl, True, False)) < 6e-3
    dh = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, Wt, dh, M, N, K, a_kmajor=True, b_kmajor=False, lda=K, ldb=N, act=ops.ACT_DSILU, preact=pre, alpha=0.5)
    h = pre.double().cpu()
    sg = torch.sigmoid(h)
    assert rel(dh, 0.5 * ref_gemm(A, Wt, True, False) * (sg * (1 + h * (1 - sg)))) < 8e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_layernorm_tail(ops, dtype):
    M, N, K = 333, 256, 512
    A, B = rnd(M, K, seed=10, dtype=dtype), rnd(N, K, seed=11, scale=0.1, dtype=dtype)
    bias, res, g, b = rnd(N, seed=12), rnd(M, N, seed=13), 1 + 0.2 * rnd(N, seed=14), rnd(N, seed=15, scale=0.1)
    x = torch.empty(M, N, device="cuda")
    u = torch.empty(M, N, device="cuda", dtype=dtype)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.gemm(A, B, x, M, N, K, bias=bias, alpha=0.5, residual=res, ln_gamma=g, ln_beta=b, ln_out=u, ln_mean=mean, ln_rstd=rstd)
    xr = 0.5 * (ref_gemm(A, B, True, True) + bias.double().cpu()) + res.double().cpu()
    mu = xr.mean(-1, keepdim=True)
    var = ((xr - mu) ** 2).mean(-1, keepdim=True)
    ur = (xr - mu) / torch.sqrt(var + 1e-5) * g.double().cpu() + b.double().cpu()
    tol = 3e-6 if dtype == torch.float32 else 1e-2
    assert rel(x, xr) < (3e-6 if dtype == torch.float32 else 1e-5)
    assert rel(u, ur) < tol
    assert rel(mean, mu.squeeze(-1)) < 1e-4 and rel(rstd, 1 / torch.sqrt(var + 1e-5).squeeze(-1)) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_logsoftmax(ops, dtype):
    M = 400
    x, w, bias = rnd(M, 256, seed=20, dtype=dtype), rnd(256, 256, seed=21, scale=0.2, dtype=dtype), rnd(256, seed=22)
    out = torch.empty(M, 256, device="cuda")
    am = torch.empty(M, dtype=torch.int32, device="cuda")
    en = torch.empty(M, device="cuda")
    ws = torch.empty(M, 256, device="cuda")
    ops.head_logsoftmax(x, w, bias, out, am, en, ws)
    logits = x.double().cpu() @ w.double().cpu().t() + bias.double().cpu()
    ref = torch.log_softmax(logits, -1)
    assert float((out.double().cpu() - ref).abs().max()) < (2e-5 if dtype == torch.float32 else 1e-4)
    assert torch.equal(am.cpu().long(), out.cpu().argmax(-1))  # consistent with its own output
    assert (am.cpu().long() == ref.argmax(-1)).float().mean() > 0.995
    assert rel(en, -(ref.exp() * ref).sum(-1)) < 1e-4


def test_layernorm_fwd_bwd(ops):
    rows = 777
    x, g, b, dy = rnd(rows, 256, seed=30, scale=2.0), 1 + 0.2 * rnd(256, seed=31), rnd(256, seed=32), rnd(rows, 256, seed=33)
    out = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, g, b, out, mean, rstd)
    xd = x.double().cpu().requires_grad_(True)
    gd, bd = g.double().cpu().requires_grad_(True), b.double().cpu().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (256,), gd, bd, 1e-5)
    assert rel(out, ref.detach()) < 2e-6
    ref.backward(dy.double().cpu())
    dx = torch.ones_like(x)
    dg, db = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, g, dx, True, dg, db)
    assert rel(dx - 1, xd.grad) < 1e-5
    assert rel(dg, gd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5
    outb = torch.empty(rows, 256, device="cuda", dtype=torch.bfloat16)
    ops.layernorm_fwd(x, g, b, outb)
    assert rel(outb, ref.detach()) < 1e-2
    # upstream gradient in bf16 (eec_layernorm_bwd_dy: what the data-gradient GEMMs of the bf16 path hand over) + bf16 copy + column sums
    dyh = dy.to(torch.bfloat16)
    xd2 = x.double().cpu().requires_grad_(True)
    gd2, bd2 = g.double().cpu().requires_grad_(True), b.double().cpu().requires_grad_(True)
    torch.nn.functional.layer_norm(xd2, (256,), gd2, bd2, 1e-5).backward(dyh.double().cpu())
    dx2 = torch.ones_like(x)
    dg2, db2, cs2 = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    copy = torch.empty(rows, 256, device="cuda", dtype=torch.bfloat16)
    ops.layernorm_bwd(dyh, x, mean, rstd, g, dx2, True, dg2, db2, copy, cs2, 0.5)
    assert rel(dx2 - 1, xd2.grad) < 1e-5
    assert rel(dg2, gd2.grad) < 1e-5 and rel(db2, bd2.grad) < 1e-5
    assert rel(copy, dx2) < 5e-3
    assert rel(cs2, 0.5 * dx2.double().cpu().sum(0)) < 1e-4     # (summed from the fp32 values, before the copy is rounded)


def attn_ref(qkv, key_len, B, T, H):
    D = H * 32
    q, k, v = qkv.double().cpu().view(B, T, 3 * D).split(D, -1)
    q, k, v = [t.view(B, T, H, 32).transpose(1, 2) for t in (q, k, v)]
    s = q @ k.transpose(-1, -2) / math.sqrt(32)
    mask = torch.arange(T)[None, :] >= key_len.cpu()[:, None].long()
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)
    return (p @ v).transpose(1, 2).reshape(B * T, D), torch.logsumexp(s, -1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T", [40, 203, 374])
def test_attention_fwd(ops, dtype, T):
    B, H = 3, 8
    qkv = rnd(B * T, 768, seed=40, dtype=dtype)
    key_len = torch.tensor([T, max(1, T // 2), 0 if T == 40 else T - 3], dtype=torch.int32, device="cuda")
    ctx = torch.empty(B * T, 256, device="cuda", dtype=dtype)
    lse = torch.empty(B, H, T, device="cuda")
    ops.attn_fwd(qkv, key_len, ctx, lse, B, T, H)
    ref, lse_ref = attn_ref(qkv, key_len, B, T, H)
    assert rel(ctx, ref) < (3e-6 if dtype == torch.float32 else 1.5e-2)
    ok = torch.isfinite(lse_ref)
    # (bf16: the tensor-core kernel's row sum is taken over the bf16-rounded probabilities the P V product actually multiplies)
    assert float((lse.double().cpu()[ok] - lse_ref[ok]).abs().max()) < (1e-5 if dtype == torch.float32 else 2e-3)
    if T == 40:  # fully masked utterance -> zeros (torch CPU SDPA behaviour)
        assert float(ctx.view(B, T, 256)[2].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_bwd(ops, dtype):
    B, T, H = 2, 150, 8
    qkv = rnd(B * T, 768, seed=41, dtype=dtype)
    dctx = rnd(B * T, 256, seed=42, dtype=dtype)
    key_len = torch.tensor([T, 97], dtype=torch.int32, device="cuda")
    ctx = torch.empty(B * T, 256, device="cuda", dtype=dtype)
    lse = torch.empty(B, H, T, device="cuda")
    ops.attn_fwd(qkv, key_len, ctx, lse, B, T, H)
    dqkv = torch.empty_like(qkv)
    dvec = torch.empty(B * H * T, device="cuda")
    ops.attn_bwd(qkv, ctx, dctx, lse, key_len, dqkv, dvec, B, T, H, torch.empty(B * T, 256, device='cuda'))
    x = qkv.double().cpu().requires_grad_(True)
    D = 256
    q, k, v = x.view(B, T, 3 * D).split(D, -1)
    q, k, v = [t.view(B, T, H, 32).transpose(1, 2) for t in (q, k, v)]
    s = q @ k.transpose(-1, -2) / math.sqrt(32)
    mask = torch.arange(T)[None, :] >= key_len.cpu()[:, None].long()
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    o.backward(dctx.double().cpu())
    assert rel(dqkv, x.grad) < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_module_interior(ops, dtype):
    B, T, C, K = 3, 150, 256, 31
    g = rnd(B, T, C, seed=50, dtype=dtype)
    w, bias = rnd(C, K, seed=51, scale=0.2), rnd(C, seed=52, scale=0.1)
    bw, bb = 1 + 0.2 * rnd(C, seed=53), rnd(C, seed=54, scale=0.1)
    rm, rv = rnd(C, seed=55, scale=0.1), torch.rand(C, generator=torch.Generator().manual_seed(56)).cuda() * 0.5 + 0.1
    gd = g.double().cpu().requires_grad_(True)
    wd = w.double().cpu().requires_grad_(True)
    bd = bias.double().cpu().requires_grad_(True)
    bwd_, bbd = bw.double().cpu().requires_grad_(True), bb.double().cpu().requires_grad_(True)
    c_ref = torch.nn.functional.conv1d(gd.transpose(1, 2), wd[:, None, :], bd, padding=15, groups=C).transpose(1, 2)
    # eval
    out = torch.empty(B, T, C, device="cuda", dtype=dtype)
    ops.dwconv_bn_silu_eval(g, w, bias, bw, bb, rm, rv, out, B, T, K)
    n = (c_ref - rm.double().cpu()) / torch.sqrt(rv.double().cpu() + 1e-5) * bwd_ + bbd
    tol = 5e-6 if dtype == torch.float32 else 1e-2
    assert rel(out, (n * torch.sigmoid(n)).detach()) < tol
    # train forward
    c = torch.empty(B * T, C, device="cuda")
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    ops.dwconv_stats(g, w, bias, c, sums, B, T, K)
    assert rel(c.view(B, T, C), c_ref.detach()) < 5e-6
    sm, sr = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    nbt = torch.tensor(3, dtype=torch.int64, device="cuda")
    s = torch.empty(B * T, C, device="cuda", dtype=dtype)
    ops.bn_silu_train(c, sums, bw, bb, rm2, rv2, nbt, 0.1, sm, sr, s)
    mean = c_ref.mean((0, 1))
    var = ((c_ref - mean) ** 2).mean((0, 1))
    nt = (c_ref - mean) / torch.sqrt(var + 1e-5) * bwd_ + bbd
    s_ref = nt * torch.sigmoid(nt)
    assert rel(s.view(B, T, C), s_ref.detach()) < tol
    Nn = B * T
    assert rel(rm2, 0.9 * rm.double().cpu() + 0.1 * mean.detach()) < 1e-5
    assert rel(rv2, 0.9 * rv.double().cpu() + 0.1 * var.detach() * Nn / (Nn - 1)) < 1e-5
    assert int(nbt) == 4
    # backward
    ds = rnd(B * T, C, seed=57, dtype=dtype)
    s_ref.backward(ds.double().cpu().view(B, T, C))
    dc = torch.empty(B * T, C, device="cuda")
    sums2 = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dgam, dbet = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn_silu_bwd(ds, c, sm, sr, bw, bb, sums2, dc, dgam, dbet)
    assert rel(dgam, bwd_.grad) < 1e-4 and rel(dbet, bbd.grad) < 1e-4
    dg = torch.empty(B * T, C, device="cuda", dtype=dtype)
    dw, dbias = torch.zeros(C, K, device="cuda"), torch.zeros(C, device="cuda")
    ops.dwconv_bwd(dc, g, w, dg, dw, dbias, B, T, K)
    btol = 2e-5 if dtype == torch.float32 else 1e-2
    assert rel(dg.view(B, T, C), gd.grad) < btol
    assert rel(dw, wd.grad) < 1e-4
    # dbias is ~0 under train-mode BN (mean removed): compare absolutely
    assert float((dbias.double().cpu() - bd.grad).abs().max()) < 1e-3
    # GLU backward
    z, dgl = rnd(B * T, 2 * C, seed=58, dtype=dtype), rnd(B * T, C, seed=59, dtype=dtype)
    dz = torch.empty_like(z)
    ops.glu_bwd(z, dgl, dz)
    zd = z.double().cpu().requires_grad_(True)
    torch.nn.functional.glu(zd, -1).backward(dgl.double().cpu())
    assert rel(dz, zd.grad) < (1e-6 if dtype == torch.float32 else 1e-2)


def test_frontend_pieces(ops):
    B, C, T_in = 2, 80, 163
    T1 = (T_in - 3) // 2 + 1
    src = rnd(B, C, T_in, seed=60)
    cols = torch.empty(B * T1, 256, device="cuda")
    ops.im2col_k3s2(src, C * T_in, T_in, 1, cols, 256, B, C, T1)
    ref = torch.stack([src[:, :, j: j + 2 * T1 - 1: 2] for j in range(3)], -1)  # (B,C,T1,3)
    ref = ref.permute(0, 2, 1, 3).reshape(B * T1, C * 3)
    assert torch.equal(cols[:, :240], ref) and float(cols[:, 240:].abs().max()) == 0
    # col2im is the adjoint of im2col on frame-major input
    D = 256
    T2 = (T1 - 3) // 2 + 1
    x1 = rnd(B * T1, D, seed=61)
    cols2 = torch.empty(B * T2, 3 * D, device="cuda")
    ops.im2col_k3s2(x1, T1 * D, 1, D, cols2, 3 * D, B, D, T2)
    ref2 = torch.stack([x1.view(B, T1, D)[:, j: j + 2 * T2 - 1: 2, :] for j in range(3)], -1).reshape(B * T2, 3 * D)   # [(b,t), c*3+j]
    assert torch.equal(cols2, ref2)
    for Tq in (163, 1501):   # bf16 outputs (what the tensor-core path consumes), ragged tile edges
        Tq1 = (Tq - 3) // 2 + 1
        Tq2 = (Tq1 - 3) // 2 + 1
        s2 = rnd(3, C, Tq, seed=63)
        cb = torch.empty(3 * Tq1, 256, device="cuda", dtype=torch.bfloat16)
        ops.im2col_k3s2(s2, C * Tq, Tq, 1, cb, 256, 3, C, Tq1)
        r = torch.stack([s2[:, :, j: j + 2 * Tq1 - 1: 2] for j in range(3)], -1).permute(0, 2, 1, 3).reshape(3 * Tq1, C * 3)
        assert torch.equal(cb[:, :240], r.to(torch.bfloat16)) and float(cb[:, 240:].float().abs().max()) == 0
        xq = rnd(3 * Tq1, D, seed=64)
        cq = torch.empty(3 * Tq2, 3 * D, device="cuda", dtype=torch.bfloat16)
        ops.im2col_k3s2(xq, Tq1 * D, 1, D, cq, 3 * D, 3, D, Tq2)
        rq = torch.stack([xq.view(3, Tq1, D)[:, j: j + 2 * Tq2 - 1: 2, :] for j in range(3)], -1).reshape(3 * Tq2, 3 * D)
        assert torch.equal(cq, rq.to(torch.bfloat16))
    dcols = rnd(B * T2, 3 * D, seed=62)
    dx = torch.empty(B * T1, D, device="cuda")
    ops.col2im_k3s2(dcols, 3 * D, dx, B, D, T1, T2)
    lhs = (cols2.double() * dcols.double()).sum()
    rhs = (x1.double() * dx.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-6 * abs(float(lhs)) + 1e-6
    kl = torch.empty(4, dtype=torch.int32, device="cuda")
    ops.encoder_lengths(torch.tensor([1500, 1499, 1000, 7], device="cuda"), kl, 374, 4, 0)
    assert kl.tolist() == [374, 374, 250, 1]


def test_misc(ops):
    x = rnd(1000, 520, seed=70)
    out = torch.ones(520, device="cuda")
    ops.colsum(x, out, 1000, 520, scale=0.5)
    assert rel(out, 1 + 0.5 * x.double().cpu().sum(0)) < 1e-5
    xb = torch.empty(1000, 520, device="cuda", dtype=torch.bfloat16)
    ops.cast(x, xb)
    assert torch.equal(xb, x.to(torch.bfloat16))
    B, T = 3, 51
    a = rnd(B * T, 256, seed=71)
    half = torch.empty(B * 26, 256, device="cuda")
    ops.stride2_gather(a, half, B, T)
    ref = torch.nn.functional.pad(a.view(B, T, 256), (0, 0, 0, 1))[:, ::2]
    assert torch.equal(half.view(B, 26, 256), ref)
    y = a.clone()
    ops.repeat2_add(half, y, B, T)
    assert torch.allclose(y.view(B, T, 256), a.view(B, T, 256) + torch.repeat_interleave(ref, 2, 1)[:, :T])


def test_wgrad_fused_bias_colsum():
    """eec_gemm_desc.a_colsum: the weight-gradient GEMM (MN-major A and B, split-K, reduce-add) also accumulates the column
    sums of its A operand (= the bias gradient) from the shared-memory tiles; compared with fp64 sums of the bf16 inputs."""
    from eec import ops
    torch.manual_seed(3)
    for rows, m, n in ((23936, 2048, 256), (1000, 768, 256), (333, 512, 256)):
        dy = (torch.randn(rows, m, device="cuda") * 0.5).to(torch.bfloat16)
        x = torch.randn(rows, n, device="cuda").to(torch.bfloat16)
        dw = torch.full((m, n), 0.25, device="cuda")          # accumulate semantics: += on top of what is there
        db = torch.full((m,), -1.0, device="cuda")
        ops.gemm(dy, x, dw, m, n, rows, a_kmajor=False, b_kmajor=False, lda=m, ldb=n, accumulate=True, a_colsum=db, a_colsum_scale=0.5)
        ref_w = 0.25 + dy.double().t() @ x.double()
        ref_b = -1.0 + 0.5 * dy.double().sum(0)
        assert float((dw.double() - ref_w).abs().max() / ref_w.abs().max()) < 2e-5
        assert float((db.double() - ref_b).abs().max() / ref_b.abs().max()) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["causal_self", "cross", "cross_long"])
def test_attention_general(ops, dtype, mode):
    """eec_attn_general_fwd / _bwd (the AED decoder's attention, early_exit.py:701-717): causal self-attention with a key-padding mask over a
    packed [rows, 768] projection, and cross-attention (Tq != Tk, Q and K/V from different tensors, K/V at a column offset of a wider
    tensor) against the same arithmetic in plain PyTorch fp32 (fp32 kernels 2e-5, bf16 tensor-core kernels 2e-2)."""
    torch.manual_seed(3)
    B, Hh, dh = 3, 8, 32
    Dm = Hh * dh
    if mode == "causal_self":
        Tq = Tk = 45
        qkv = (torch.randn(B * Tq, 3 * Dm, device="cuda") * 0.7).to(dtype)
        q, k, v = qkv[:, :Dm], qkv[:, Dm:2 * Dm], qkv[:, 2 * Dm:]
        toks = torch.randint(3, 120, (B, Tk), device="cuda")
        toks[1, 30:] = 126
        toks[2, 7:] = 126
        toks[2, 3] = 126                  # a padded key in the middle
        bits = ops.key_bits_from_tokens(toks, 126)
        valid = toks != 126
        causal = True
    else:
        Tq, Tk = (37, 200) if mode == "cross" else (130, 390)
        qt = (torch.randn(B * Tq, Dm, device="cuda") * 0.7).to(dtype)
        kvw = (torch.randn(B * Tk, 4 * Dm, device="cuda") * 0.7).to(dtype)     # K | V of "layer 1" of a stacked projection
        q, k, v = qt, kvw[:, 2 * Dm:3 * Dm], kvw[:, 3 * Dm:]
        bits, valid, causal = None, torch.ones(B, Tk, dtype=torch.bool, device="cuda"), False
    ctx = torch.empty(B * Tq, Dm, device="cuda", dtype=dtype)
    lse = torch.empty(B, Hh, Tq, device="cuda")
    ops.attn_general_fwd(q, k, v, ctx, lse, B, Tq, Tk, Hh, key_bits=bits, causal=causal)
    # reference
    qf = q.float().reshape(B, Tq, Hh, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
    kf = k.float().reshape(B, Tk, Hh, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
    vf = v.float().reshape(B, Tk, Hh, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
    sc = qf @ kf.transpose(-1, -2) / dh ** 0.5
    mask = valid[:, None, None, :].expand(B, Hh, Tq, Tk).clone()
    if causal:
        mask &= torch.tril(torch.ones(Tq, Tk, dtype=torch.bool, device="cuda"))[None, None]
    sc = sc.masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(sc, -1) @ vf)
    ref_ctx = ref.permute(0, 2, 1, 3).reshape(B * Tq, Dm)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert rel(ctx, ref_ctx) < tol
    assert rel(lse, torch.logsumexp(sc, -1)) < (1e-5 if dtype == torch.float32 else 2e-2)
    dctx = (torch.randn(B * Tq, Dm, device="cuda") * 0.5).to(dtype)
    ref_ctx.backward(dctx.float())
    dq = torch.empty(B * Tq, Dm, device="cuda", dtype=dtype)
    dkv = torch.zeros(B * Tk, 2 * Dm, device="cuda", dtype=dtype)
    ops.attn_general_bwd(q, k, v, ctx, dctx, lse, dq, dkv[:, :Dm], dkv[:, Dm:], B, Tq, Tk, Hh, key_bits=bits, causal=causal)
    btol = 1e-4 if dtype == torch.float32 else 3e-2
    assert rel(dq, qf.grad.permute(0, 2, 1, 3).reshape(B * Tq, Dm)) < btol
    assert rel(dkv[:, :Dm], kf.grad.permute(0, 2, 1, 3).reshape(B * Tk, Dm)) < btol
    assert rel(dkv[:, Dm:], vf.grad.permute(0, 2, 1, 3).reshape(B * Tk, Dm)) < btol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_relu_epilogues(ops, dtype):
    """RELU / DRELU epilogues of eec_gemm (nn.TransformerDecoderLayer's feed-forward and its backward)"""
    torch.manual_seed(5)
    M, N, K = 300, 2048, 256
    a = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.gemm(a, w, out, M, N, K, bias=bias, act=ops.ACT_RELU)
    ref = torch.relu(a.float() @ w.float().t() + bias)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert rel(out, ref) < tol
    dy = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)       # dY [M, K_in=256] @ W2 [256, 2048] -> [M, 2048], times [a > 0]
    w2 = (torch.randn(K, N, device="cuda") * 0.05).to(dtype)
    dh = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.gemm(dy, w2, dh, M, N, K, a_kmajor=True, b_kmajor=False, lda=K, ldb=N, act=ops.ACT_DRELU, preact=out, alpha=0.5)
    refd = 0.5 * (dy.float() @ w2.float()) * (out.float() > 0)
    assert rel(dh, refd) < tol
