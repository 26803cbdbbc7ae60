"""Forward / backward of the AED decoder stacks over libeec.so kernels.

Reference arithmetic: ``full_conformer``'s six ``nn.TransformerDecoder`` stacks (early_exit.py:701-717: ``TransformerDecoderLayer(d_model 256,
nhead 8, dim_feedforward 2048, batch_first, norm_first)`` x n_dec_layers with ONE shared final ``layer_norm``), called at :772-800 on
``emb(trg) + pe`` with the causal ``tgt_mask`` and ``tgt_key_padding_mask = (trg == pad)``, NO memory mask (padded encoder frames are
attended to), followed by ``linears_2`` (raw logits; the log-softmax is commented out at :790).  Per pre-norm layer (torch
nn/modules/transformer.py, norm_first branch):

    y = y + SelfAttn(LN1(y))   causal + key padding        y = y + CrossAttn(LN2(y), memory)        y = y + W2 relu(W1 LN3(y) + b1) + b2

Like ``eec.engine`` this is a manual tape: torch tensors are device buffers, every contraction is an ``eec_gemm`` (tcgen05 in bf16 mode;
LayerNorm tails and the ReLU / dReLU epilogues fused), the attention cores are ``eec_attn_general_fwd / _bwd``.  The K / V projections of
the encoder states for ALL layers of a stack are ONE GEMM (their weights stacked: [n_dec * 512, 256]), and so are their weight and
data gradients.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .engine import D, F, H, V, Config, Operands, _empty, dgrad, linear, to_act, wgrad
from .lib import ACT_DRELU, ACT_RELU, EecError

Tensor = torch.Tensor
# dropout sites (train mode, drop_prob > 0; include/eec.h "dropout"): after emb + positional encoding (positional_encoding.py:72); per
# decoder layer (stack e, layer l) SITE_BASE + 8 * (e * n_dec + l) + k -- nn.TransformerDecoderLayer's self-attention probabilities
# (nn.MultiheadAttention(dropout = p)) and dropout1, cross-attention probabilities and dropout2, the feed-forward's inner dropout and dropout3
SITE_PE, SITE_BASE = 99999, 100000
D_SA_P, D_SA_OUT, D_CA_P, D_CA_OUT, D_FF_ACT, D_FF_OUT = range(6)


def _layer_prefix(e: int, l: int) -> str:
    return f"decoders.{e}.layers.{l}."


def decoder_param_names(n_exits: int, n_dec: int) -> List[str]:
    """decoder-side parameters in the order named_parameters() yields them (the shared final norm appears once, as `layer_norm.*`)"""
    names = ["layer_norm.weight", "layer_norm.bias", "emb.weight"]
    names += [f"linears_2.{e}.{s}" for e in range(n_exits) for s in ("weight", "bias")]
    for e in range(n_exits):
        for l in range(n_dec):
            p = _layer_prefix(e, l)
            names += [p + s for s in ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                                      "multihead_attn.in_proj_weight", "multihead_attn.in_proj_bias", "multihead_attn.out_proj.weight",
                                      "multihead_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                                      "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "norm3.weight", "norm3.bias")]
    return names


def _kv_stack(P, W: Operands, e: int, n_dec: int, cfg: Config):
    """[n_dec * 512, 256] operand + [n_dec * 512] bias: rows 256..767 (K | V) of every layer's multihead_attn.in_proj of stack e"""
    w = torch.cat([P[_layer_prefix(e, l) + "multihead_attn.in_proj_weight"].detach()[D:] for l in range(n_dec)], 0)      # layout plumbing
    b = torch.cat([P[_layer_prefix(e, l) + "multihead_attn.in_proj_bias"].detach()[D:] for l in range(n_dec)], 0)
    return to_act(w.contiguous(), cfg), b.contiguous()


def decoder_forward(P: Dict[str, Tensor], W: Operands, cfg: Config, n_dec: int, trg: Tensor, pad_idx: int, hidden: Tensor,
                    exits: List[int], mem_index: List[int], want_tape: bool, drop0: Optional[ops.Drop] = None):
    """trg [B, L] int64 (device); hidden [M, B, T, D] fp32 encoder states; exits: which decoder stacks to run, stack exits[i] attending to
    hidden[mem_index[i]].  -> (logits [len(exits), B, L, V] fp32, tape | None)"""
    if not hidden.is_cuda or not trg.is_cuda:
        raise EecError("eec decoder: tensors must be on a CUDA device; there is no CPU path")
    B, Ln = trg.shape
    T = hidden.shape[2]
    Nd, Ne = B * Ln, B * T
    dev, TD, f32 = hidden.device, cfg.act_dtype, torch.float32
    pe = P["positional_encoder_2.pe"]
    if Ln > pe.shape[0]:
        raise RuntimeError(f"The size of tensor a ({Ln}) must match the size of tensor b ({pe.shape[0]}) at non-singleton dimension 0")
    trg = trg.contiguous()
    x0 = _empty((Nd, D), f32, dev)
    ops.embed_pe(trg, P["emb.weight"].detach(), pe.view(-1, D), x0)           # early_exit.py:776-777 (no sqrt(d) scaling)
    bf16 = cfg.precision == "bf16"
    if drop0 is not None:
        ops.dropout(x0, x0, drop0.at(SITE_PE))                                # positional_encoding.py:72

    def site(e, l, k, R=0, C=0, Cs=0, Wd=0):
        """dropout site k of decoder layer (e, l); the tensor-core kernels get its keep-mask words (generated here, shared with backward)"""
        if drop0 is None:
            return None
        d = drop0.at(SITE_BASE + 8 * (e * n_dec + l) + k)
        return d.with_bits(R, C, Cs, Wd) if (bf16 and Wd) else d
    key_bits = ops.key_bits_from_tokens(trg, pad_idx)                          # tgt_key_padding_mask (:773-775, :802-805)
    out = _empty((len(exits), B, Ln, V), f32, dev)
    tape = {"B": B, "L": Ln, "T": T, "trg": trg, "key_bits": key_bits, "exits": [], "n_dec": n_dec, "drop0": drop0} if want_tape else None
    L8, T8 = 8 * ((Ln + 7) // 8), 8 * ((T + 7) // 8)

    def stat():
        return (_empty((Nd,), f32, dev), _empty((Nd,), f32, dev)) if want_tape else (None, None)

    for oi, e in enumerate(exits):
        mem = to_act(hidden[mem_index[oi]].reshape(Ne, D), cfg)
        wkv, bkv = _kv_stack(P, W, e, n_dec, cfg)
        NK = n_dec * 2 * D
        kvall = _empty((Ne, NK), TD, dev)
        for c0 in range(0, NK, 2048):             # (the GEMM stages at most 2048 bias values per launch: 4 layers' K | V at a time)
            c1 = min(c0 + 2048, NK)
            linear(mem, wkv[c0:c1], kvall[:, c0:c1], Ne, c1 - c0, D, bias=bkv[c0:c1], ldc=NK)
        et = {"e": e, "mi": mem_index[oi], "mem": mem, "wkv": wkv, "kvall": kvall, "layers": []} if want_tape else None
        y = x0
        p0 = _layer_prefix(e, 0)
        u = _empty((Nd, D), TD, dev)
        m1, r1 = stat()
        ops.layernorm_fwd(y, P[p0 + "norm1.weight"], P[p0 + "norm1.bias"], u, m1, r1)
        for l in range(n_dec):
            p = _layer_prefix(e, l)
            nxt_g, nxt_b = ((P[_layer_prefix(e, l + 1) + "norm1.weight"], P[_layer_prefix(e, l + 1) + "norm1.bias"]) if l + 1 < n_dec
                            else (P["layer_norm.weight"], P["layer_norm.bias"]))
            # ---- self-attention (causal + key padding)
            Wsi = W.get(p + "self_attn.in_proj_weight", P[p + "self_attn.in_proj_weight"], (3 * D, D))
            Wso = W.get(p + "self_attn.out_proj.weight", P[p + "self_attn.out_proj.weight"], (D, D))
            qkv = _empty((Nd, 3 * D), TD, dev)
            linear(u, Wsi, qkv, Nd, 3 * D, D, bias=P[p + "self_attn.in_proj_bias"])
            ctx = _empty((Nd, D), TD, dev)
            lse1 = _empty((B, H, Ln), f32, dev)
            d_sap = site(e, l, D_SA_P, B * H * Ln, Ln, L8, 32)
            ops.attn_general_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], ctx, lse1, B, Ln, Ln, H, key_bits=key_bits, causal=True, drop=d_sap)
            y1 = _empty((Nd, D), f32, dev)
            u2 = _empty((Nd, D), TD, dev)
            m2, r2 = stat()
            linear(ctx, Wso, y1, Nd, D, D, bias=P[p + "self_attn.out_proj.bias"], residual=y, ln_gamma=P[p + "norm2.weight"],
                   ln_beta=P[p + "norm2.bias"], ln_out=u2, ln_mean=m2, ln_rstd=r2, drop=site(e, l, D_SA_OUT, Nd, D, D, 32))
            # ---- cross-attention over the encoder states (no memory mask)
            Wci = W.get(p + "multihead_attn.in_proj_weight", P[p + "multihead_attn.in_proj_weight"], (3 * D, D))
            Wco = W.get(p + "multihead_attn.out_proj.weight", P[p + "multihead_attn.out_proj.weight"], (D, D))
            q2 = _empty((Nd, D), TD, dev)
            linear(u2, Wci[:D], q2, Nd, D, D, bias=P[p + "multihead_attn.in_proj_bias"].detach()[:D])
            ctx2 = _empty((Nd, D), TD, dev)
            lse2 = _empty((B, H, Ln), f32, dev)
            c0 = l * 2 * D
            d_cap = site(e, l, D_CA_P, B * H * Ln, T, T8, 32)
            ops.attn_general_fwd(q2, kvall[:, c0:c0 + D], kvall[:, c0 + D:c0 + 2 * D], ctx2, lse2, B, Ln, T, H, drop=d_cap)
            y2 = _empty((Nd, D), f32, dev)
            u3 = _empty((Nd, D), TD, dev)
            m3, r3 = stat()
            linear(ctx2, Wco, y2, Nd, D, D, bias=P[p + "multihead_attn.out_proj.bias"], residual=y1, ln_gamma=P[p + "norm3.weight"],
                   ln_beta=P[p + "norm3.bias"], ln_out=u3, ln_mean=m3, ln_rstd=r3, drop=site(e, l, D_CA_OUT, Nd, D, D, 32))
            # ---- feed-forward (ReLU)
            W1 = W.get(p + "linear1.weight", P[p + "linear1.weight"], (F, D))
            W2 = W.get(p + "linear2.weight", P[p + "linear2.weight"], (D, F))
            a = _empty((Nd, F), TD, dev)
            d_act = site(e, l, D_FF_ACT, Nd, F, F, 16)
            linear(u3, W1, a, Nd, F, D, bias=P[p + "linear1.bias"], act=ACT_RELU, drop=d_act)
            y3 = _empty((Nd, D), f32, dev)
            un = _empty((Nd, D), TD, dev)
            mn, rn = stat()
            linear(a, W2, y3, Nd, D, F, bias=P[p + "linear2.bias"], residual=y2, ln_gamma=nxt_g, ln_beta=nxt_b, ln_out=un, ln_mean=mn,
                   ln_rstd=rn, drop=site(e, l, D_FF_OUT, Nd, D, D, 32))
            if want_tape:
                et["layers"].append(dict(y=y, u=u, m1=m1, r1=r1, qkv=qkv, ctx=ctx, lse1=lse1, y1=y1, u2=u2, m2=m2, r2=r2, q2=q2, ctx2=ctx2,
                                         lse2=lse2, y2=y2, u3=u3, m3=m3, r3=r3, a=a, y3=y3, mn=mn, rn=rn, d_sap=d_sap, d_cap=d_cap, d_act=d_act))
            y, u, m1, r1 = y3, un, mn, rn
        Wl = W.get(f"linears_2.{e}.weight", P[f"linears_2.{e}.weight"], (V, D))
        linear(u, Wl, out[oi].view(Nd, V), Nd, V, D, bias=P[f"linears_2.{e}.bias"])
        if want_tape:
            et["z"] = u
            tape["exits"].append(et)
    return out, tape


def decoder_backward(P, W: Operands, cfg: Config, tape: dict, gout: Tensor, names: List[str], ghid: Tensor):
    """gout: grad wrt the logits [len(exits), B, L, V] fp32.  Adds the encoder-state gradients into ghid [E, B, T, D] (fp32, zero-initialised by
    the caller) and returns fp32 grads for every decoder parameter name in `names` (views of one flat buffer)."""
    dev, f32, TD = gout.device, torch.float32, cfg.act_dtype
    bf16 = cfg.precision == "bf16"
    B, Ln, T, n_dec = tape["B"], tape["L"], tape["T"], tape["n_dec"]
    Nd, Ne = B * Ln, B * T
    total = sum(P[n].numel() for n in names)
    flat = torch.zeros(total, dtype=f32, device=dev)
    G, off = {}, 0
    for n in names:
        k = P[n].numel()
        G[n] = flat[off:off + k].view(P[n].shape)
        off += k
    G["__flat__"] = flat
    gout = gout.contiguous()
    dX0 = torch.zeros(Nd, D, dtype=f32, device=dev)
    drop0 = tape.get("drop0")

    def osite(e, l, k):
        return drop0.at(SITE_BASE + 8 * (e * n_dec + l) + k) if drop0 is not None else None

    for oi, et in enumerate(tape["exits"]):
        e = et["e"]
        dX = _empty((Nd, D), f32, dev)

        def ln_bwd(dy, x_in, m, r, key, accumulate, bias_key=None, want_h=True, dsite=None):
            """LayerNorm backward into the residual-gradient stream dX (+ operand copy of the new dX, + its column sums = the bias gradient
            of the projection whose output joined the residual stream right before this LayerNorm in the forward pass; dsite: that output
            went through dropout -- copy and column sums carry dropout'(dX), the residual gradient itself does not)"""
            dXh = _empty((Nd, D), TD, dev) if ((bf16 or dsite is not None) and want_h) else None
            ops.layernorm_bwd(dy, x_in, m, r, P[key + "weight"], dX, accumulate, G[key + "weight"], G[key + "bias"], dXh,
                              G[bias_key] if bias_key else None, 1.0, drop=dsite)
            return dXh if dXh is not None else dX

        # ---- linears_2 + the shared final LayerNorm
        dlog = gout[oi].view(Nd, V)
        Wl = W.get(f"linears_2.{e}.weight", P[f"linears_2.{e}.weight"], (V, D))
        if bf16:
            dlogh = _empty((Nd, V), torch.bfloat16, dev)
            ops.cast_colsum(dlog, dlogh, G[f"linears_2.{e}.bias"], Nd, V)
        else:
            dlogh = dlog
            ops.colsum(dlog, G[f"linears_2.{e}.bias"], Nd, V)
        wgrad(dlogh, et["z"], G[f"linears_2.{e}.weight"], Nd, V, D)
        dz = _empty((Nd, D), f32, dev)
        dgrad(dlogh, Wl, dz, Nd, V, D)
        last = et["layers"][-1]
        dXh = ln_bwd(dz, last["y3"], last["mn"], last["rn"], "layer_norm.", False, _layer_prefix(e, n_dec - 1) + "linear2.bias",
                     dsite=osite(e, n_dec - 1, D_FF_OUT))
        dkvall = _empty((Ne, n_dec * 2 * D), TD, dev)
        for l in reversed(range(n_dec)):
            p = _layer_prefix(e, l)
            t = et["layers"][l]
            # feed-forward: y3 = y2 + W2 relu(W1 u3 + b1) + b2
            W1 = W.get(p + "linear1.weight", P[p + "linear1.weight"], (F, D))
            W2 = W.get(p + "linear2.weight", P[p + "linear2.weight"], (D, F))
            wgrad(dXh, t["a"], G[p + "linear2.weight"], Nd, D, F)
            dh = _empty((Nd, F), TD, dev)
            dgrad(dXh, W2, dh, Nd, D, F, act=ACT_DRELU, preact=t["a"], drop=t["d_act"])
            wgrad(dh, t["u3"], G[p + "linear1.weight"], Nd, F, D, dbias=G[p + "linear1.bias"])
            du3 = _empty((Nd, D), f32, dev)
            dgrad(dh, W1, du3, Nd, F, D)
            dXh = ln_bwd(du3, t["y2"], t["m3"], t["r3"], p + "norm3.", True, p + "multihead_attn.out_proj.bias", dsite=osite(e, l, D_CA_OUT))
            # cross-attention: y2 = y1 + Wo ctx2 + bo
            Wci = W.get(p + "multihead_attn.in_proj_weight", P[p + "multihead_attn.in_proj_weight"], (3 * D, D))
            Wco = W.get(p + "multihead_attn.out_proj.weight", P[p + "multihead_attn.out_proj.weight"], (D, D))
            wgrad(dXh, t["ctx2"], G[p + "multihead_attn.out_proj.weight"], Nd, D, D)
            dctx2 = _empty((Nd, D), TD, dev)
            dgrad(dXh, Wco, dctx2, Nd, D, D)
            dq2 = _empty((Nd, D), TD, dev)
            c0 = l * 2 * D
            kv = et["kvall"]
            ops.attn_general_bwd(t["q2"], kv[:, c0:c0 + D], kv[:, c0 + D:c0 + 2 * D], t["ctx2"], dctx2, t["lse2"], dq2,
                                 dkvall[:, c0:c0 + D], dkvall[:, c0 + D:c0 + 2 * D], B, Ln, T, H, drop=t["d_cap"])
            wgrad(dq2, t["u2"], G[p + "multihead_attn.in_proj_weight"][:D], Nd, D, D, dbias=G[p + "multihead_attn.in_proj_bias"][:D])
            du2 = _empty((Nd, D), f32, dev)
            dgrad(dq2, Wci[:D], du2, Nd, D, D)
            dXh = ln_bwd(du2, t["y1"], t["m2"], t["r2"], p + "norm2.", True, p + "self_attn.out_proj.bias", dsite=osite(e, l, D_SA_OUT))
            # self-attention: y1 = y + Wo ctx + bo
            Wsi = W.get(p + "self_attn.in_proj_weight", P[p + "self_attn.in_proj_weight"], (3 * D, D))
            Wso = W.get(p + "self_attn.out_proj.weight", P[p + "self_attn.out_proj.weight"], (D, D))
            wgrad(dXh, t["ctx"], G[p + "self_attn.out_proj.weight"], Nd, D, D)
            dctx = _empty((Nd, D), TD, dev)
            dgrad(dXh, Wso, dctx, Nd, D, D)
            dqkv = _empty((Nd, 3 * D), TD, dev)
            qkv = t["qkv"]
            ops.attn_general_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], t["ctx"], dctx, t["lse1"], dqkv[:, :D], dqkv[:, D:2 * D],
                                 dqkv[:, 2 * D:], B, Ln, Ln, H, key_bits=tape["key_bits"], causal=True, drop=t["d_sap"])
            wgrad(dqkv, t["u"], G[p + "self_attn.in_proj_weight"], Nd, 3 * D, D, dbias=G[p + "self_attn.in_proj_bias"])
            du = _empty((Nd, D), f32, dev)
            dgrad(dqkv, Wsi, du, Nd, 3 * D, D)
            prev_bias = _layer_prefix(e, l - 1) + "linear2.bias" if l > 0 else None
            dXh = ln_bwd(du, t["y"], t["m1"], t["r1"], p + "norm1.", True, prev_bias, want_h=l > 0,
                         dsite=osite(e, l - 1, D_FF_OUT) if l > 0 else None)
        ops.axpy(dX, 1.0, dX0)                                   # every stack reads the same embedded targets
        # ---- K / V projections of the encoder states, all layers of the stack at once
        dwkv = torch.zeros(n_dec * 2 * D, D, dtype=f32, device=dev)
        dbkv = torch.zeros(n_dec * 2 * D, dtype=f32, device=dev)
        wgrad(dkvall, et["mem"], dwkv, Ne, n_dec * 2 * D, D, dbias=dbkv)
        for l in range(n_dec):                                   # un-stack into the parameters' own gradient slices (layout plumbing)
            p = _layer_prefix(e, l)
            G[p + "multihead_attn.in_proj_weight"][D:].copy_(dwkv[l * 2 * D:(l + 1) * 2 * D])
            G[p + "multihead_attn.in_proj_bias"][D:].copy_(dbkv[l * 2 * D:(l + 1) * 2 * D])
        gh = ghid[et["mi"]].view(Ne, D)
        dgrad(dkvall, et["wkv"], gh, Ne, n_dec * 2 * D, D, residual=gh)
    if drop0 is not None:
        ops.dropout(dX0, dX0, drop0.at(SITE_PE))
    ops.embed_bwd(tape["trg"], dX0, G["emb.weight"])
    return G
