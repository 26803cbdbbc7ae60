"""CPU oracle for the early-exit conformer hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The shipped path (``early-exit-transformer_b200/eec``) never
imports anything from ``oracle/`` and fails loudly when ``libeec.so`` is missing.

It is a plain-PyTorch (CPU, fp32 or fp64) *restatement* of the arithmetic the
reference executes for ``Early_conformer`` / ``Splitformer`` with CTC heads.  It
does not import the reference nor ``torchaudio``; every function cites the
reference lines it follows.  Citation convention (same as SURVEY.md):

  ``early_exit.py:L``  -> /root/reference/models/model/early_exit.py
  ``pos_enc.py:L``     -> /root/reference/models/embedding/positional_encoding.py
  ``train.py:L``       -> /root/reference/train.py
  ``beam_infer.py:L``  -> /root/reference/util/beam_infer.py
  ``TA:L``             -> torchaudio 2.11.0 ``torchaudio/models/conformer.py``
                          (third-party, un-vendored dependency of the reference;
                          call sites early_exit.py:16, :603-615, :627)

Parity pin: ``oracle/make_golden.py`` runs the *real* reference module (imported
from /root/reference, with torchaudio's Conformer and ``torch.nn.CTCLoss``) on
seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.
The reference itself ships no tests or golden vectors for this path,
so the pin is "outputs of the reference run here", not a reference KAT.
The early-exit selection (`early_exit_select`) has no reference implementation
at all: **parity unpinned** for that function.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- #
# deterministic parameters with the reference's state_dict key layout
# --------------------------------------------------------------------------- #
def layer_param_shapes(d: int, f: int, k: int) -> Dict[str, Tuple[int, ...]]:
    """Per-ConformerLayer tensors, keyed by their state_dict suffix (TA:122-174)."""
    s: Dict[str, Tuple[int, ...]] = {}
    for ffn in ("ffn1", "ffn2"):
        s[f"{ffn}.sequential.0.weight"] = (d,)
        s[f"{ffn}.sequential.0.bias"] = (d,)
        s[f"{ffn}.sequential.1.weight"] = (f, d)
        s[f"{ffn}.sequential.1.bias"] = (f,)
        s[f"{ffn}.sequential.4.weight"] = (d, f)
        s[f"{ffn}.sequential.4.bias"] = (d,)
    s["self_attn_layer_norm.weight"] = (d,)
    s["self_attn_layer_norm.bias"] = (d,)
    s["self_attn.in_proj_weight"] = (3 * d, d)
    s["self_attn.in_proj_bias"] = (3 * d,)
    s["self_attn.out_proj.weight"] = (d, d)
    s["self_attn.out_proj.bias"] = (d,)
    s["conv_module.layer_norm.weight"] = (d,)
    s["conv_module.layer_norm.bias"] = (d,)
    s["conv_module.sequential.0.weight"] = (2 * d, d, 1)
    s["conv_module.sequential.0.bias"] = (2 * d,)
    s["conv_module.sequential.2.weight"] = (d, 1, k)
    s["conv_module.sequential.2.bias"] = (d,)
    s["conv_module.sequential.3.weight"] = (d,)
    s["conv_module.sequential.3.bias"] = (d,)
    s["conv_module.sequential.3.running_mean"] = (d,)
    s["conv_module.sequential.3.running_var"] = (d,)
    s["conv_module.sequential.3.num_batches_tracked"] = ()
    s["conv_module.sequential.5.weight"] = (d, d, 1)
    s["conv_module.sequential.5.bias"] = (d,)
    s["final_layer_norm.weight"] = (d,)
    s["final_layer_norm.bias"] = (d,)
    return s


def positional_table(max_len: int, d: int, dtype=torch.float32) -> Tensor:
    """Sinusoid buffer ``pe (max_len,1,d)`` exactly as pos_enc.py:59-64 builds it."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2) * (-math.log(10000.0) / d))
    pe = torch.zeros(max_len, 1, d)
    pe[:, 0, 0::2] = torch.sin(position * div_term)
    pe[:, 0, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def make_params(
    seed: int,
    n_exits: int = 6,
    n_layers: int = 2,
    d: int = 256,
    f: int = 2048,
    k: int = 31,
    n_mels: int = 80,
    vocab: int = 256,
    max_len: int = 2000,
    splitformer: bool = False,
) -> Dict[str, Tensor]:
    """Deterministic random parameters under the reference's 413-key state_dict layout
    (early_exit.py:594-615).  Matrices are Xavier-uniform-like (what
    util/model_utils.py:10-12 applies), and -- unlike the reference's defaults -- every
    1-D parameter and BatchNorm buffer is randomised so that gamma/beta/running-stat
    bugs are visible (SURVEY §4 "test-data hygiene")."""
    g = torch.Generator().manual_seed(seed)

    def U(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def mat(shape):
        fan_out = shape[0] * (shape[2] if len(shape) == 3 else 1)
        fan_in = shape[1] * (shape[2] if len(shape) == 3 else 1)
        return U(shape, math.sqrt(6.0 / (fan_in + fan_out)))

    sd: Dict[str, Tensor] = {}
    sd["conv_subsample.sequential.0.weight"] = mat((d, n_mels, 3))
    sd["conv_subsample.sequential.0.bias"] = U((d,), 0.05)
    sd["conv_subsample.sequential.1.weight"] = mat((d, d, 3))
    sd["conv_subsample.sequential.1.bias"] = U((d,), 0.05)
    sd["positional_encoder.pe"] = positional_table(max_len, d)
    for e in range(n_exits):
        sd[f"linears.{e}.weight"] = mat((vocab, d))
        sd[f"linears.{e}.bias"] = U((vocab,), 0.05)

    def fill_layer(prefix: str):
        for suffix, shape in layer_param_shapes(d, f, k).items():
            key = prefix + suffix
            if suffix.endswith("num_batches_tracked"):
                sd[key] = torch.tensor(3, dtype=torch.int64)
            elif suffix.endswith("running_var"):
                sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.05
            elif suffix.endswith("running_mean"):
                sd[key] = U(shape, 0.05)
            elif len(shape) >= 2:
                sd[key] = mat(shape)
            elif suffix.endswith(".weight") and ("norm" in suffix or ".0.weight" in suffix or ".3.weight" in suffix):
                sd[key] = 1.0 + U(shape, 0.2)  # LN / BN gamma
            else:
                sd[key] = U(shape, 0.05)  # biases, LN / BN beta

    for e in range(n_exits):
        for l in range(n_layers):
            fill_layer(f"conformer.{e}.conformer_layers.{l}.")
    if splitformer:
        for i in range(2):
            fill_layer(f"conformer_parallel.{i}.conformer_layers.0.")
    return sd


# --------------------------------------------------------------------------- #
# forward restatement
# --------------------------------------------------------------------------- #
def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """LayerNorm over the last dim, biased variance, eps 1e-5 (TA:103,151,42,211)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def silu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def conv_subsample(src: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """Two Conv1d(k=3,s=2) with NO activation in between (early_exit.py:24-48).
    src (B,n_mels,T_in) -> (B,T',d) frame-major."""
    w1, b1 = sd["conv_subsample.sequential.0.weight"], sd["conv_subsample.sequential.0.bias"]
    w2, b2 = sd["conv_subsample.sequential.1.weight"], sd["conv_subsample.sequential.1.bias"]
    # written as explicit unfold + matmul so the arithmetic is visible
    def conv_k3s2(x, w, b):  # x (B,C,T)
        T = (x.shape[2] - 3) // 2 + 1
        cols = torch.stack([x[:, :, j : j + 2 * T - 1 : 2] for j in range(3)], dim=-1)  # (B,C,T,3)
        return torch.einsum("bctj,ocj->bot", cols, w) + b[None, :, None]
    x2 = conv_k3s2(conv_k3s2(src, w1, b1), w2, b2)
    return x2.permute(0, 2, 1)


def encoder_lengths(lengths: Tensor, t_out: int) -> Tensor:
    """early_exit.py:623: clamp(lengths / 4, max=T').to(int) -- true divide, truncation."""
    return torch.clamp(lengths / 4, max=t_out).to(torch.int)


# Dropout sites (train mode, drop_prob > 0).  The reference draws its masks from torch's RNG, which no other implementation
# can reproduce; the oracle therefore takes the masks as an INPUT: `drop(site, shape, row_stride=None)` returns the
# keep-mask already multiplied by 1/(1-p) (oracle/philox_dropout.Masks restates the product's counter-based generator).
# Site ids: 0 = after the positional encoding (pos_enc.py:72); layer uid u (main stack: e*L + l, Splitformer branch i:
# 1000 + i) owns sites 8*(u+1) + k with k = the S_* constants below (the nn.Dropout modules of TA:106/108, TA:152, TA:201, TA:73).
S_FFN1_ACT, S_FFN1_OUT, S_ATTN_P, S_ATTN_OUT, S_CONV_OUT, S_FFN2_ACT, S_FFN2_OUT = range(7)


def _drop(x: Tensor, drop, site: Optional[int], row_stride: Optional[int] = None) -> Tensor:
    if drop is None or site is None:
        return x
    return x * torch.as_tensor(drop(site, tuple(x.shape), row_stride)).to(x.dtype)


def ffn(x: Tensor, sd, p: str, drop=None, s_act: Optional[int] = None, s_out: Optional[int] = None) -> Tensor:
    """TA:102-109: LN -> Linear(d,F) -> SiLU -> Dropout -> Linear(F,d) -> Dropout."""
    u = layer_norm(x, sd[p + "sequential.0.weight"], sd[p + "sequential.0.bias"])
    h = u @ sd[p + "sequential.1.weight"].T + sd[p + "sequential.1.bias"]
    a = _drop(silu(h), drop, s_act)
    return _drop(a @ sd[p + "sequential.4.weight"].T + sd[p + "sequential.4.bias"], drop, s_out)


def mhsa(x: Tensor, key_len: Tensor, sd, p: str, n_head: int, drop=None, s_p: Optional[int] = None,
         s_out: Optional[int] = None) -> Tensor:
    """TA:192-202 -> nn.MultiheadAttention: packed in-proj, per-head softmax(QK^T/sqrt(dh)
    + key-padding mask) V, out-proj.  Keys t' >= key_len[b] get -inf; a fully masked row
    yields 0 (torch 2.11 CPU SDPA behaviour)."""
    B, T, D = x.shape
    dh = D // n_head
    u = layer_norm(x, sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
    qkv = u @ sd[p + "self_attn.in_proj_weight"].T + sd[p + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, T, n_head, dh).transpose(1, 2)
    k = k.view(B, T, n_head, dh).transpose(1, 2)
    v = v.view(B, T, n_head, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)  # (B,H,T,T)
    masked = torch.arange(T)[None, :] >= key_len[:, None].to(torch.int64)  # (B,T)
    s = s.masked_fill(masked[:, None, None, :], float("-inf"))
    m = s.max(-1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m)
    den = e.sum(-1, keepdim=True)
    pr = torch.where(den > 0, e / torch.where(den > 0, den, torch.ones_like(den)), torch.zeros_like(e))
    pr = _drop(pr, drop, s_p, row_stride=8 * ((T + 7) // 8))  # MultiheadAttention(dropout=p): on the probabilities (TA:152)
    o = (pr @ v).transpose(1, 2).reshape(B, T, D)
    return _drop(o @ sd[p + "self_attn.out_proj.weight"].T + sd[p + "self_attn.out_proj.bias"], drop, s_out)  # TA:201


def conv_module(x: Tensor, sd, p: str, training: bool, bn_out: Optional[dict], drop=None, s_out: Optional[int] = None) -> Tensor:
    """TA:42-75, 85-88: LN -> pw conv d->2d -> GLU -> depthwise conv k (SAME, zero pad per
    utterance) -> BatchNorm1d (train: batch stats over all B*T frames incl. padding;
    eval: running stats) -> SiLU -> pw conv d->d."""
    q = p + "conv_module."
    B, T, D = x.shape
    u = layer_norm(x, sd[q + "layer_norm.weight"], sd[q + "layer_norm.bias"])
    z = u @ sd[q + "sequential.0.weight"][:, :, 0].T + sd[q + "sequential.0.bias"]
    g = z[..., :D] * torch.sigmoid(z[..., D:])
    wd = sd[q + "sequential.2.weight"][:, 0, :]  # (D,K)
    K = wd.shape[1]
    half = (K - 1) // 2
    gp = F.pad(g, (0, 0, half, half))  # zero pad time
    c = sd[q + "sequential.2.bias"] + sum(gp[:, j : j + T, :] * wd[:, j] for j in range(K))
    gam, bet = sd[q + "sequential.3.weight"], sd[q + "sequential.3.bias"]
    if training:
        mean = c.mean(dim=(0, 1))
        var = ((c - mean) ** 2).mean(dim=(0, 1))
        if bn_out is not None:
            n = B * T
            rm, rv = sd[q + "sequential.3.running_mean"], sd[q + "sequential.3.running_var"]
            bn_out[q + "sequential.3.running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
            bn_out[q + "sequential.3.running_var"] = ((1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / (n - 1)).detach()
            bn_out[q + "sequential.3.num_batches_tracked"] = sd[q + "sequential.3.num_batches_tracked"] + 1
    else:
        mean, var = sd[q + "sequential.3.running_mean"], sd[q + "sequential.3.running_var"]
    nrm = (c - mean) / torch.sqrt(var + BN_EPS) * gam + bet
    return _drop(silu(nrm) @ sd[q + "sequential.5.weight"][:, :, 0].T + sd[q + "sequential.5.bias"], drop, s_out)  # TA:73


def conformer_layer(x, key_len, sd, p, n_head, training=False, bn_out=None, drop=None, uid: int = 0) -> Tensor:
    """TA:176-212 with convolution_first=False: FFN/2 -> MHSA -> conv -> FFN/2 -> LN."""
    base = 8 * (uid + 1)
    x = x + 0.5 * ffn(x, sd, p + "ffn1.", drop, base + S_FFN1_ACT, base + S_FFN1_OUT)
    x = x + mhsa(x, key_len, sd, p, n_head, drop, base + S_ATTN_P, base + S_ATTN_OUT)
    x = x + conv_module(x, sd, p, training, bn_out, drop, base + S_CONV_OUT)
    x = x + 0.5 * ffn(x, sd, p + "ffn2.", drop, base + S_FFN2_ACT, base + S_FFN2_OUT)
    return layer_norm(x, sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])


def _count_layers(sd, prefix: str) -> int:
    n = 0
    while f"{prefix}conformer_layers.{n}.ffn1.sequential.0.weight" in sd:
        n += 1
    return n


def early_conformer_forward(
    sd: Dict[str, Tensor],
    src: Tensor,
    lengths: Tensor,
    n_head: int = 8,
    training: bool = False,
    bn_out: Optional[dict] = None,
    splitformer: bool = False,
    return_hidden: bool = False,
    drop=None,
):
    """early_exit.py:617-634 (Early_conformer) / :299-364 (Splitformer).
    src (B,n_mels,T_in), lengths (B,) int64 fbank frame counts -> (E,B,T',V) log-probs."""
    x = conv_subsample(src, sd)
    B, T, D = x.shape
    x = x + sd["positional_encoder.pe"][:T, 0, :].to(x.dtype)  # pos_enc.py:70-72
    if not training:
        drop = None
    x = _drop(x, drop, 0)  # pos_enc.py:72
    key_len = encoder_lengths(lengths, T)
    if int(key_len.max()) < T:
        # TA:11-14 builds a mask of width max(length); nn.MultiheadAttention then asserts
        raise AssertionError(f"Expected key_padded_mask.shape[1] to be {T}, but got {int(key_len.max())}")
    n_exits = 0
    while f"linears.{n_exits}.weight" in sd:
        n_exits += 1
    outs, hidden = [], []
    for e in range(n_exits):
        x_in = x
        n_l = _count_layers(sd, f"conformer.{e}.")
        for l in range(n_l):
            x = conformer_layer(x, key_len, sd, f"conformer.{e}.conformer_layers.{l}.", n_head, training, bn_out, drop, e * n_l + l)
        if splitformer and (e == 0 or e == n_exits - 1):
            i = e // (n_exits - 1)
            pad = T % 2
            xd = F.pad(x_in, (0, 0, 0, pad)) if pad else x_in  # early_exit.py:318-327
            xd = xd[:, ::2, :]  # :329-331
            len2 = torch.clamp((lengths + pad) / 2, max=xd.shape[1]).to(torch.int)  # :332-338 (raw lengths!)
            xd = conformer_layer(xd, len2, sd, f"conformer_parallel.{i}.conformer_layers.0.", n_head, training, bn_out, drop, 1000 + i)
            xd = torch.repeat_interleave(xd, 2, dim=1)  # :344-346
            if pad:
                xd = xd[:, :-pad, :]
            x = x + xd  # :356
        hidden.append(x)
        logits = x @ sd[f"linears.{e}.weight"].T + sd[f"linears.{e}.bias"]
        outs.append(torch.log_softmax(logits, dim=2))  # early_exit.py:629-631
    out = torch.stack(outs, 0)
    return (out, hidden) if return_hidden else out


# --------------------------------------------------------------------------- #
# CTC (train.py:57-65, 259: nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True))
# --------------------------------------------------------------------------- #
def ctc_alpha_beta(lp: Tensor, target: Tensor, blank: int = 0):
    """Log-space alpha/beta recursions for ONE utterance.  lp (T,V) float64/32,
    target (U,) ints.  Returns (nll, grad_wrt_logits (T,V)) with
    grad = exp(lp) - occupancy  (Graves eq. 16), unscaled."""
    T, V = lp.shape
    U = int(target.numel())
    S = 2 * U + 1
    ext = torch.full((S,), blank, dtype=torch.int64)
    ext[1::2] = target.to(torch.int64)
    NEG = float("-inf")
    la = torch.full((T, S), NEG, dtype=lp.dtype)
    lb = torch.full((T, S), NEG, dtype=lp.dtype)
    la[0, 0] = lp[0, blank]
    if S > 1:
        la[0, 1] = lp[0, ext[1]]
    can_skip = torch.zeros(S, dtype=torch.bool)
    can_skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    for t in range(1, T):
        prev = la[t - 1]
        a1 = torch.full((S,), NEG, dtype=lp.dtype)
        a1[1:] = prev[:-1]
        a2 = torch.full((S,), NEG, dtype=lp.dtype)
        a2[2:] = prev[:-2]
        a2 = torch.where(can_skip, a2, torch.full_like(a2, NEG))
        la[t] = torch.logsumexp(torch.stack([prev, a1, a2]), 0) + lp[t, ext]
    lb[T - 1, S - 1] = lp[T - 1, blank]
    if S > 1:
        lb[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    skip_fwd = torch.zeros(S, dtype=torch.bool)
    skip_fwd[:-2] = can_skip[2:]
    for t in range(T - 2, -1, -1):
        nxt = lb[t + 1]
        b1 = torch.full((S,), NEG, dtype=lp.dtype)
        b1[:-1] = nxt[1:]
        b2 = torch.full((S,), NEG, dtype=lp.dtype)
        b2[: max(S - 2, 0)] = nxt[2:]
        b2 = torch.where(skip_fwd, b2, torch.full_like(b2, NEG))
        lb[t] = torch.logsumexp(torch.stack([nxt, b1, b2]), 0) + lp[t, ext]
    tail = la[T - 1, S - 1 :] if S == 1 else la[T - 1, S - 2 :]
    ll = torch.logsumexp(tail, 0)
    nll = -ll
    if torch.isinf(nll):
        return nll, torch.zeros_like(lp)
    occ_log = la + lb  # alpha*beta includes lp[t,s] twice
    acc = torch.full((T, V), NEG, dtype=lp.dtype)
    for s in range(S):
        c = int(ext[s])
        acc[:, c] = torch.logaddexp(acc[:, c], occ_log[:, s])
    grad = torch.exp(lp) - torch.exp(acc - lp - ll)
    return nll, grad


def ctc_loss_mean(lp_btv: Tensor, targets: Tensor, target_lengths: Tensor, blank: int = 0):
    """One exit's nn.CTCLoss(reduction='mean', zero_infinity=True) with input length = T'
    for every utterance (train.py:57-58).  lp_btv (B,T,V).  Returns (loss, dlogits (B,T,V))."""
    B = lp_btv.shape[0]
    loss = lp_btv.new_zeros(())
    grads = torch.zeros_like(lp_btv)
    for b in range(B):
        U = int(target_lengths[b])
        nll, g = ctc_alpha_beta(lp_btv[b], targets[b, :U], blank)
        if torch.isinf(nll):  # zero_infinity
            continue
        denom = max(U, 1) * B
        loss = loss + nll / denom
        grads[b] = g / denom
    return loss, grads


def multi_exit_ctc(out_ebtv: Tensor, targets: Tensor, target_lengths: Tensor):
    """train.py:60-63: sum over exits of the mean CTC loss.  Returns (loss, per_exit, dlogits)."""
    per, grads = [], []
    for e in range(out_ebtv.shape[0]):
        l, g = ctc_loss_mean(out_ebtv[e], targets, target_lengths)
        per.append(l)
        grads.append(g)
    per_t = torch.stack(per)
    return per_t.sum(), per_t, torch.stack(grads)


# --------------------------------------------------------------------------- #
# decode + early exit
# --------------------------------------------------------------------------- #
def greedy_ctc(lp_tv: Tensor, blank: int = 0) -> List[int]:
    """beam_infer.py:21-23: argmax -> unique_consecutive -> drop blank, over ALL given frames."""
    idx = torch.argmax(lp_tv, dim=-1).tolist()
    out, prev = [], None
    for i in idx:
        if i != prev:
            if i != blank:
                out.append(i)
            prev = i
    return out


def frame_entropy_mean(lp_btv: Tensor, key_len: Tensor) -> Tensor:
    """Mean over valid frames t < key_len[b] of -sum_c p log p.  NOT in the
    reference: parity unpinned."""
    p = lp_btv.exp()
    h = -(torch.where(p > 0, p * lp_btv, torch.zeros_like(p))).sum(-1)  # (B,T)
    T = lp_btv.shape[1]
    valid = torch.arange(T)[None, :] < key_len[:, None].to(torch.int64)
    n = valid.sum(1).clamp(min=1)
    return (h * valid).sum(1) / n


def early_exit_select(out_ebtv: Tensor, key_len: Tensor, threshold: float):
    """exit_b = first e with mean frame entropy < threshold, else last exit; tokens = greedy
    of that exit.  NOT in the reference: parity unpinned."""
    E, B = out_ebtv.shape[:2]
    H = torch.stack([frame_entropy_mean(out_ebtv[e], key_len) for e in range(E)])  # (E,B)
    exit_idx = torch.full((B,), E - 1, dtype=torch.int64)
    for b in range(B):
        for e in range(E):
            if H[e, b] < threshold:
                exit_idx[b] = e
                break
    tokens = [greedy_ctc(out_ebtv[int(exit_idx[b]), b]) for b in range(B)]
    return exit_idx, tokens, H


# --------------------------------------------------------------------------- #
# synthetic workload of SURVEY §8(d)
# --------------------------------------------------------------------------- #
def synthetic_batch(B: int, t_in: int, seed: int = 1234, n_mels: int = 80, min_frac: float = 0.5):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(B, n_mels, t_in, generator=g)
    lengths = torch.randint(int(t_in * min_frac), t_in + 1, (B,), generator=g)
    lengths[0] = t_in
    for b in range(B):
        src[b, :, int(lengths[b]) :] = 0.0
    return src, lengths.to(torch.int64)


def synthetic_targets(B: int, seed: int = 4321, lo: int = 20, hi: int = 80):
    """Rows [<s>=1, tokens in 3..125, </s>=2, pad=126...] (util/data_loader.py:207-214)."""
    g = torch.Generator().manual_seed(seed)
    tl = torch.randint(lo, hi + 1, (B,), generator=g)
    L = int(tl.max()) + 2
    tg = torch.full((B, L), 126, dtype=torch.int64)
    for b in range(B):
        n = int(tl[b])
        tg[b, 0] = 1
        tg[b, 1 : 1 + n] = torch.randint(3, 126, (n,), generator=g)
        tg[b, 1 + n] = 2
    return tg, (tl + 2).to(torch.int64)
