"""Thin Python wrappers over the C ABI: one function per kernel family.  Tensors are torch CUDA
tensors used purely as device-memory handles (data_ptr + shape); all arithmetic happens in
libeec.so on torch's current stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L
from .lib import ACT_DRELU, ACT_DSILU, ACT_GLU, ACT_NONE, ACT_RELU, ACT_SILU, BF16, F32, call, dt, ptr, stream

Tensor = torch.Tensor


class Drop:
    """One dropout site of one forward pass: (device state {seed, offset} as an int64[2] tensor, probability, site id).
    The keep mask is a pure function of these three and the element index (include/eec.h "dropout").  `bits` (optional) holds
    the mask as 1 bit per element (eec_dropout_bits) for the tensor-core kernels, whose epilogues do not run Philox themselves;
    the same words serve the forward and the backward kernel of the site."""
    __slots__ = ("state", "p", "site", "bits")

    def __init__(self, state: Tensor, p: float, site: int, bits: Optional[Tensor] = None):
        self.state, self.p, self.site, self.bits = state, float(p), int(site), bits

    def at(self, k: int) -> "Drop":
        return Drop(self.state, self.p, self.site + k)

    def alloc_bits(self, R: int, C: int, Cs: int, W: int):
        """-> (this site with an UNFILLED word buffer, fill()): allocation and generation are separate so that the engine can
        allocate on the compute stream and generate on a side stream (logical tensor [R, C], element index r*Cs + c)"""
        words = R * ((C + W - 1) // W)
        bits = torch.empty(words, dtype=torch.int16 if W == 16 else torch.int32, device=self.state.device)
        d = Drop(self.state, self.p, self.site, bits)

        def fill():
            call("eec_dropout_bits", d.state.data_ptr(), d.p, d.site, R, C, Cs, W, bits.data_ptr(), stream())
        return d, fill

    def with_bits(self, R: int, C: int, Cs: int, W: int) -> "Drop":
        """this site with its keep-mask words generated (one launch on the current stream)"""
        if self.p <= 0.0:
            return self
        d, fill = self.alloc_bits(R, C, Cs, W)
        fill()
        return d


def _d(drop):
    """(state pointer, p, site) C arguments of an optional Drop"""
    if drop is None or drop.p <= 0.0:
        return None, 0.0, 0
    return drop.state.data_ptr(), drop.p, drop.site


def _chk(t: Tensor, name: str):
    if not t.is_cuda:
        raise L.EecError(f"{name}: tensor must live on a CUDA device (no CPU path exists)")
    if not t.is_contiguous():
        raise L.EecError(f"{name}: tensor must be contiguous")


def gemm(
    A: Tensor, B: Tensor, C_out: Tensor, M: int, N: int, K: int, *,
    a_kmajor: bool = True, b_kmajor: bool = True, lda: Optional[int] = None, ldb: Optional[int] = None,
    bias: Optional[Tensor] = None, act: int = ACT_NONE, preact: Optional[Tensor] = None, alpha: float = 1.0,
    residual: Optional[Tensor] = None, res_row_mod: int = 0, accumulate: bool = False,
    ln_gamma: Optional[Tensor] = None, ln_beta: Optional[Tensor] = None, ln_out: Optional[Tensor] = None,
    ln_mean: Optional[Tensor] = None, ln_rstd: Optional[Tensor] = None, ldc: Optional[int] = None,
    a_colsum: Optional[Tensor] = None, a_colsum_scale: float = 1.0, drop: Optional[Drop] = None,
):
    """C[M,N] = epi(A(m,k) B(n,k)).  See include/eec.h::eec_gemm_desc."""
    for t, n in ((A, "A"), (B, "B"), (C_out, "C")):
        if n == "C" and ldc is not None and t.is_cuda and t.dim() == 2 and t.stride(1) == 1:
            continue          # a column block of a wider row-major tensor: the caller passes its row pitch as ldc
        _chk(t, "gemm." + n)
    if A.dtype != B.dtype:
        raise L.EecError("gemm: A and B dtypes differ")
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_kmajor = ptr(A), (lda if lda is not None else (K if a_kmajor else M)), int(a_kmajor)
    d.B, d.ldb, d.b_kmajor = ptr(B), (ldb if ldb is not None else (K if b_kmajor else N)), int(b_kmajor)
    d.in_dtype = dt(A)
    d.bias = ptr(bias)
    d.act = act
    d.preact = ptr(preact)
    d.ldp = N
    d.preact_dtype = dt(preact) if preact is not None else F32
    d.alpha = alpha
    d.residual = ptr(residual)
    d.ldr = N
    d.res_row_mod = res_row_mod
    d.C = ptr(C_out)
    n_out = N // 2 if act == ACT_GLU else N
    d.ldc = ldc if ldc is not None else n_out
    d.out_dtype = dt(C_out)
    d.accumulate = int(accumulate)
    d.ln_gamma, d.ln_beta, d.ln_out = ptr(ln_gamma), ptr(ln_beta), ptr(ln_out)
    d.ln_dtype = dt(ln_out) if ln_out is not None else F32
    d.ld_ln = N
    d.ln_mean, d.ln_rstd = ptr(ln_mean), ptr(ln_rstd)
    d.a_colsum, d.a_colsum_scale = ptr(a_colsum), a_colsum_scale
    if drop is not None and drop.p > 0.0 and drop.bits is None and A.dtype == torch.bfloat16:
        drop = drop.with_bits(M, N, N, 32 if ln_out is not None else 16)    # (callers that also run the backward keep the words)
    d.drop_state, d.drop_p, d.drop_site = _d(drop)
    d.drop_bits = ptr(drop.bits) if (drop is not None and drop.p > 0.0) else None
    call("eec_gemm", C.byref(d), stream())


def ffn_fwd(u, w1, b1, w2, b2, residual, alpha, ln_gamma, ln_beta, x_out, ln_out, ln_mean=None, ln_rstd=None, hpre=None):
    """Fused feed-forward module + following LayerNorm (include/eec_experiments.h::eec_ffn_fwd, libeec_exp.so only); bf16 operands."""
    for t, n in ((u, "u"), (w1, "w1"), (w2, "w2"), (residual, "residual"), (x_out, "x_out"), (ln_out, "ln_out")):
        _chk(t, "ffn_fwd." + n)
    rows, f = u.numel() // 256, w1.shape[0]
    call("eec_ffn_fwd", ptr(u), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(residual), alpha, ptr(ln_gamma), ptr(ln_beta),
         ptr(x_out), ptr(ln_out), dt(ln_out), ptr(ln_mean), ptr(ln_rstd), ptr(hpre), rows, 256, f, stream())


def layernorm_fwd(x, gamma, beta, out, mean=None, rstd=None):
    rows = x.numel() // 256
    call("eec_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(out), dt(out), ptr(mean), ptr(rstd), rows, 256, stream())


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, accumulate, dgamma, dbeta, dx_copy=None, dx_colsum=None, colsum_scale=1.0,
                  drop: Optional[Drop] = None):
    rows = x.numel() // 256
    call("eec_layernorm_bwd_dy", ptr(dy), dt(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dx), int(accumulate), ptr(dgamma),
         ptr(dbeta), ptr(dx_copy), dt(dx_copy) if dx_copy is not None else BF16, ptr(dx_colsum), colsum_scale, *_d(drop), rows, 256,
         stream())


def attn_drop(drop: Optional[Drop], qkv, B, T, H) -> Optional[Drop]:
    """the attention-probability site with keep-mask words when the tensor-core kernels will run (bf16 operands)"""
    if drop is None or drop.p <= 0.0 or drop.bits is not None or qkv.dtype != torch.bfloat16:
        return drop
    return drop.with_bits(B * H * T, T, 8 * ((T + 7) // 8), 32)


def _bits(drop):
    return ptr(drop.bits) if (drop is not None and drop.p > 0.0) else None


def attn_fwd(qkv, key_len, ctx, lse, B, T, H, drop: Optional[Drop] = None):
    drop = attn_drop(drop, qkv, B, T, H)
    call("eec_attn_fwd", ptr(qkv), dt(qkv), ptr(key_len), ptr(ctx), ptr(lse), B, T, H, 32, *_d(drop), _bits(drop), stream())


def attn_bwd(qkv, ctx, dctx, lse, key_len, dqkv, dvec, B, T, H, dq32=None, drop: Optional[Drop] = None):
    drop = attn_drop(drop, qkv, B, T, H)
    call("eec_attn_bwd", ptr(qkv), ptr(ctx), ptr(dctx), dt(qkv), ptr(lse), ptr(key_len), ptr(dqkv), ptr(dvec), ptr(dq32), B, T,
         H, 32, *_d(drop), _bits(drop), stream())


def _attn_desc(q, k, v, B, Tq, Tk, H, key_len, key_bits, causal, drop=None):
    """q / k / v: 2-D views (rows, H*32) of row-major tensors -- a column block of a packed projection output is passed as its slice"""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
            raise L.EecError(f"attn.{n}: expected a 2-D CUDA view with unit column stride")
    if not (q.dtype == k.dtype == v.dtype):
        raise L.EecError("attn: q / k / v dtypes differ")
    d = L.AttnDesc()
    d.B, d.H, d.dh, d.Tq, d.Tk = B, H, 32, Tq, Tk
    d.q, d.ldq = q.data_ptr(), q.stride(0)
    d.k, d.ldk = k.data_ptr(), k.stride(0)
    d.v, d.ldv = v.data_ptr(), v.stride(0)
    d.dtype = dt(q)
    d.key_len, d.key_valid_bits, d.causal = ptr(key_len), ptr(key_bits), int(bool(causal))
    if drop is not None and drop.p > 0.0:
        if drop.bits is None and q.dtype == torch.bfloat16:
            raise L.EecError("attn: bf16 dropout needs the site's keep-mask words (Drop.with_bits(B*H*Tq, Tk, 8*ceil(Tk/8), 32))")
        d.drop_state, d.drop_p, d.drop_site = _d(drop)
        d.drop_bits = ptr(drop.bits)
    return d


def attn_general_fwd(q, k, v, ctx, lse, B, Tq, Tk, H, key_len=None, key_bits=None, causal=False, drop=None):
    """include/eec.h::eec_attn_general_fwd (decoder self-attention with causal / padding masks, cross-attention over encoder states)"""
    d = _attn_desc(q, k, v, B, Tq, Tk, H, key_len, key_bits, causal, drop)
    call("eec_attn_general_fwd", C.byref(d), ptr(ctx), ctx.stride(0), ptr(lse), stream())


def attn_general_bwd(q, k, v, ctx, dctx, lse, dq, dk, dv, B, Tq, Tk, H, key_len=None, key_bits=None, causal=False, drop=None):
    """dq / dk / dv: 2-D views like q / k / v (written, not accumulated)"""
    d = _attn_desc(q, k, v, B, Tq, Tk, H, key_len, key_bits, causal, drop)
    dev = q.device
    dvec = torch.empty(B * H * Tq, dtype=torch.float32, device=dev)
    dq32 = torch.empty(B * Tq, H * 32, dtype=torch.float32, device=dev) if q.dtype == torch.bfloat16 else None
    call("eec_attn_general_bwd", C.byref(d), ptr(ctx), ptr(dctx), ctx.stride(0), ptr(lse), dq.data_ptr(), dq.stride(0), dk.data_ptr(),
         dk.stride(0), dv.data_ptr(), dv.stride(0), ptr(dvec), ptr(dq32), stream())


def key_bits_from_tokens(tokens, pad):
    """uint32 words [B, ceil(L/32)]: bit set iff tokens[b, t] != pad"""
    B, Ln = tokens.shape
    bits = torch.empty(B, (Ln + 31) // 32, dtype=torch.int32, device=tokens.device)
    call("eec_key_bits_from_tokens", ptr(tokens), B, Ln, int(pad), ptr(bits), stream())
    return bits


def embed_pe(tokens, emb, pe, x):
    B, Ln = tokens.shape
    call("eec_embed_pe", ptr(tokens), ptr(emb), ptr(pe), ptr(x), B, Ln, emb.shape[1], emb.shape[0], stream())


def embed_bwd(tokens, dx, demb):
    B, Ln = tokens.shape
    call("eec_embed_bwd", ptr(tokens), ptr(dx), ptr(demb), B, Ln, demb.shape[1], demb.shape[0], stream())


def cross_entropy(logits, targets, loss_out, dlogits=None):
    rows, V = logits.shape
    call("eec_cross_entropy", ptr(logits), ptr(targets), rows, V, ptr(loss_out), ptr(dlogits), stream())


def dropout(x, y, drop: Drop):
    """y = x * keep_mask * 1/(1-p) for the dropout site `drop` (x, y contiguous, same numel; in place allowed)."""
    _chk(x, "dropout.x")
    _chk(y, "dropout.y")
    call("eec_dropout", ptr(x), dt(x), ptr(y), dt(y), x.numel(), drop.state.data_ptr(), drop.p, drop.site, stream())


def dropout_advance(state):
    call("eec_dropout_advance", ptr(state), stream())


def dropout_p_effective(p: float) -> float:
    """the probability the kernels actually use: p quantised to 1/65536 (include/eec.h "dropout")"""
    return min(int(p * 65536.0 + 0.5), 65535) / 65536.0


def dwconv_bn_silu_eval(g, w, bias, bn_w, bn_b, rm, rv, out, B, T, K):
    call("eec_dwconv_bn_silu_eval", ptr(g), dt(g), ptr(w), ptr(bias), ptr(bn_w), ptr(bn_b), ptr(rm), ptr(rv), ptr(out), B, T,
         256, K, stream())


def dwconv_stats(g, w, bias, c, sums, B, T, K):
    call("eec_dwconv_stats", ptr(g), dt(g), ptr(w), ptr(bias), ptr(c), ptr(sums), B, T, 256, K, stream())


def bn_silu_train(c, sums, bn_w, bn_b, rm, rv, nbt, momentum, save_mean, save_rstd, out, stat_rows=0):
    rows = c.numel() // 256
    call("eec_bn_silu_train", ptr(c), ptr(sums), ptr(bn_w), ptr(bn_b), ptr(rm), ptr(rv), ptr(nbt), momentum, ptr(save_mean),
         ptr(save_rstd), ptr(out), dt(out), rows, 256, stat_rows, stream())


def bn_silu_bwd(ds, c, save_mean, save_rstd, bn_w, bn_b, sums2, dc, dgamma, dbeta, sync=None, world=1):
    """sync (optional): in-place SUM all-reduce of a double tensor over the data-parallel ranks (synchronised BatchNorm)"""
    rows = c.numel() // 256
    call("eec_bn_silu_bwd_stats", ptr(ds), dt(ds), ptr(c), ptr(save_mean), ptr(save_rstd), ptr(bn_w), ptr(bn_b), ptr(sums2),
         rows, 256, stream())
    local = None
    if sync is not None:
        local = sums2.clone()
        sync(sums2)
    call("eec_bn_silu_bwd_apply", ptr(ds), dt(ds), ptr(c), ptr(save_mean), ptr(save_rstd), ptr(bn_w), ptr(bn_b), ptr(sums2),
         ptr(dc), ptr(dgamma), ptr(dbeta), rows, 256, rows * world if sync is not None else 0, ptr(local), stream())


def dwconv_bwd(dc, g, w, dg, dw, dbias, B, T, K):
    ws = torch.empty(L.load().eec_dwconv_bwd_workspace_bytes(B, T, 256) // 4, dtype=torch.float32, device=dc.device)
    call("eec_dwconv_bwd", ptr(dc), ptr(g), dt(g), ptr(w), ptr(dg), ptr(dw), ptr(dbias), B, T, 256, K, ptr(ws), stream())


def glu_bwd(z, dg, dz):
    rows = dg.numel() // 256
    call("eec_glu_bwd", ptr(z), ptr(dg), ptr(dz), dt(z), rows, 256, stream())


def head_logsoftmax(x, w, bias, out, argmax=None, entropy=None, logits_ws=None):
    rows = x.numel() // 256
    call("eec_head_logsoftmax", ptr(x), dt(x), ptr(w), ptr(bias), ptr(out), ptr(argmax), ptr(entropy), ptr(logits_ws), rows,
         256, 256, stream())


def logsoftmax_bwd(g, lp, dlogits):
    rows = lp.numel() // 256
    call("eec_logsoftmax_bwd", ptr(g), ptr(lp), ptr(dlogits), rows, 256, stream())


def ctc_fwd_bwd(lp, targets, target_len, nll, loss_out, grad, gscale=1.0, blank=0):
    """lp [E,B,T,V] fp32; targets [B,Lmax] int64 (device); target_len [B] int64 (device)."""
    E, B, T, V = lp.shape
    Lmax = targets.shape[1]
    nbytes = L.load().eec_ctc_workspace_bytes(E, B, T, Lmax)
    ws = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=lp.device)
    call("eec_ctc_fwd_bwd", ptr(lp), ptr(targets), ptr(target_len), E, B, T, V, Lmax, blank, gscale, ptr(nll), ptr(loss_out),
         ptr(grad), ptr(ws), stream())


def greedy_collapse(argmax, tokens, n_tokens, B, T, blank=0):
    call("eec_greedy_collapse", ptr(argmax), ptr(tokens), ptr(n_tokens), B, T, blank, stream())


def ctc_beam_search(lp, enc_len, beam, nbest, blank, log_blank_skip):
    """lp [..., T, V] fp32 log-probs (contiguous) -> (tokens [n_utt, nbest, T] int32 (-1 padded), n_tokens [n_utt, nbest], scores [n_utt, nbest])"""
    _chk(lp, "ctc_beam_search.lp")
    T, V = lp.shape[-2], lp.shape[-1]
    n_utt = lp.numel() // (T * V)
    dev = lp.device
    tokens = torch.empty(n_utt, nbest, T, dtype=torch.int32, device=dev)
    n_tok = torch.empty(n_utt, nbest, dtype=torch.int32, device=dev)
    scores = torch.empty(n_utt, nbest, dtype=torch.float32, device=dev)
    ws = torch.empty(max(L.load().eec_ctc_beam_workspace_bytes(n_utt, T, beam), 4) // 4, dtype=torch.int32, device=dev)
    call("eec_ctc_beam_search", ptr(lp), ptr(enc_len), n_utt, T, V, beam, nbest, blank, log_blank_skip, ptr(tokens), ptr(n_tok),
         ptr(scores), ptr(ws), stream())
    return tokens, n_tok, scores


def im2col_k3s2(inp, sb, sc, st, out, ldo, B, Cin, T_out):
    call("eec_im2col_k3s2", ptr(inp), dt(inp), sb, sc, st, ptr(out), dt(out), ldo, B, Cin, T_out, stream())


def col2im_k3s2(dcols, ldc, dx, B, Cc, T_in, T_out):
    call("eec_col2im_k3s2", ptr(dcols), ldc, ptr(dx), B, Cc, T_in, T_out, stream())


def encoder_lengths(lengths_dev, key_len, T, div=4, add=0):
    call("eec_encoder_lengths", ptr(lengths_dev), ptr(key_len), lengths_dev.numel(), T, div, add, stream())


def cast(inp, out):
    call("eec_cast", ptr(inp), dt(inp), ptr(out), dt(out), inp.numel(), stream())


def colsum(inp, out, rows, cols, scale=1.0, ld=None):
    call("eec_colsum", ptr(inp), dt(inp), ld if ld is not None else cols, ptr(out), scale, rows, cols, stream())


def cast_colsum(inp, out_bf16, colsum_out, rows, cols, scale=1.0):
    call("eec_cast_colsum", ptr(inp), cols, ptr(out_bf16), cols, ptr(colsum_out), scale, rows, cols, stream())


def axpy(x, a, y):
    call("eec_axpy", ptr(x), a, ptr(y), x.numel(), stream())


def scale_rows_dev(x, s_dev, y):
    """y[r] = s_dev[r] * x[r] for the leading dimension r (contiguous tensors)."""
    call("eec_scale_rows_dev", ptr(x), ptr(s_dev), ptr(y), x.shape[0], x.numel() // x.shape[0], stream())


def scale_dev(x, s_dev, y):
    call("eec_scale_dev", ptr(x), ptr(s_dev), ptr(y), x.numel(), stream())


def exit_select(entropy, argmax, key_len_alive, row_map, n_alive, exit_idx, is_last, threshold, exit_index, tokens, n_tokens,
                new_row_map, new_key_len, gather_idx, mean_entropy, B, T, blank=0):
    call("eec_exit_select", ptr(entropy), ptr(argmax), ptr(key_len_alive), ptr(row_map), ptr(n_alive), exit_idx, int(is_last),
         threshold, ptr(exit_index), ptr(tokens), ptr(n_tokens), ptr(new_row_map), ptr(new_key_len), ptr(gather_idx),
         ptr(mean_entropy), B, T, blank, stream())


def gather_rows(x, y, gather_idx, n_alive, B, row_elems):
    call("eec_gather_rows", ptr(x), ptr(y), ptr(gather_idx), ptr(n_alive), B, row_elems, stream())


def set_active_items(n_items_dev, rows_per_item: int = 0, pad_items: int = 0):
    """Limit the inference kernels launched on the current stream to the first *n_items_dev (+ pad_items) utterances (None clears)."""
    call("eec_set_active_items", ptr(n_items_dev), int(rows_per_item), int(pad_items), stream())


def gather_i64(src, idx, dst):
    call("eec_gather_i64", ptr(src), ptr(idx), ptr(dst), src.numel(), stream())


def stride2_gather(x, y, B, T, D=256):
    call("eec_stride2_gather", ptr(x), ptr(y), B, T, D, stream())


def repeat2_add(up, y, B, T, D=256):
    call("eec_repeat2_add", ptr(up), ptr(y), B, T, D, stream())


def repeat2_bwd(dy, dhalf, B, T, D=256):
    call("eec_repeat2_bwd", ptr(dy), ptr(dhalf), B, T, D, stream())


def stride2_scatter_add(dhalf, dx, B, T, D=256):
    call("eec_stride2_scatter_add", ptr(dhalf), ptr(dx), B, T, D, stream())
