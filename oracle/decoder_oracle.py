"""CPU oracle of the AED decoder half of ``full_conformer``.  TEST INFRASTRUCTURE ONLY (imported by tests/ only; the
product path, early-exit-transformer_b200/eec/decoder_engine.py, never touches it).

Restates, in plain PyTorch tensor arithmetic without torch.nn modules, what the reference executes at early_exit.py:772-800 (forward),
:739-762 (``_decoder_``) and train.py:36-51 (the AED loss): ``emb(trg)`` + ``positional_encoder_2`` (positional_encoding.py:65-73, no
sqrt(d) scaling), then per exit a stack of ``nn.TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, batch_first, norm_first)``
(torch 2.11 nn/modules/transformer.py, norm_first branch of ``forward``: ``x = x + _sa_block(norm1(x))``, ``x = x + _mha_block(norm2(x), memory)``,
``x = x + _ff_block(norm3(x))`` with ReLU), the ONE shared final LayerNorm (early_exit.py:666, :713) and ``linears_2`` (raw logits).
Masks: the causal ``tgt_mask`` plus ``tgt_key_padding_mask = (trg == pad)``; the encoder states are NOT masked (no memory mask is passed).

Parity pin: tests/golden/fc_*.npz -- outputs of the real reference (oracle/make_golden.py::run_fc_case) -- checked in
tests/test_oracle_golden.py::test_decoder_oracle_vs_reference_golden.  Dropout follows oracle/conformer_oracle.py's convention: the masks
are an input (``drop(site, shape, row_stride)``, oracle/philox_dropout.Masks); site ids are those of eec.decoder_engine.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from .conformer_oracle import _drop, layer_norm

Tensor = torch.Tensor
SITE_PE = 99999                      # dropout after emb + positional encoding (positional_encoding.py:72)
SITE_BASE = 100000                   # decoder layer (stack e, layer l) owns sites SITE_BASE + 8 * (e * n_dec + l) + k:
D_SA_P, D_SA_OUT, D_CA_P, D_CA_OUT, D_FF_ACT, D_FF_OUT = range(6)


def _mha(q_in: Tensor, kv_in: Tensor, w: Tensor, b: Tensor, wo: Tensor, bo: Tensor, n_head: int, mask: Optional[Tensor], drop, s_p) -> Tensor:
    """nn.MultiheadAttention with packed in_proj (rows 0-255 q, 256-511 k, 512-767 v); mask (B, Tq, Tk) True = visible, or None."""
    B, Tq, D = q_in.shape
    Tk = kv_in.shape[1]
    dh = D // n_head
    q = (q_in @ w[:D].T + b[:D]).view(B, Tq, n_head, dh).transpose(1, 2)
    k = (kv_in @ w[D:2 * D].T + b[D:2 * D]).view(B, Tk, n_head, dh).transpose(1, 2)
    v = (kv_in @ w[2 * D:].T + b[2 * D:]).view(B, Tk, n_head, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if mask is not None:
        s = s.masked_fill(~mask[:, None], float("-inf"))
    pr = torch.softmax(s, dim=-1)
    pr = _drop(pr, drop, s_p, row_stride=8 * ((Tk + 7) // 8))
    o = (pr @ v).transpose(1, 2).reshape(B, Tq, D)
    return o @ wo.T + bo


def decoder_stack(sd: Dict[str, Tensor], e: int, n_dec: int, x: Tensor, memory: Tensor, mask: Tensor, n_head: int = 8, drop=None) -> Tensor:
    """decoders[e] of the reference on the embedded targets x (B, L, D) and the encoder state memory (B, T', D) -> (B, L, D) after the
    shared final LayerNorm."""
    for l in range(n_dec):
        p = f"decoders.{e}.layers.{l}."
        base = SITE_BASE + 8 * (e * n_dec + l)
        u = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = x + _drop(_mha(u, u, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"], sd[p + "self_attn.out_proj.weight"],
                           sd[p + "self_attn.out_proj.bias"], n_head, mask, drop, base + D_SA_P), drop, base + D_SA_OUT)
        u = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        x = x + _drop(_mha(u, memory, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                           sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], n_head, None, drop, base + D_CA_P),
                      drop, base + D_CA_OUT)
        u = layer_norm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        a = _drop(torch.relu(u @ sd[p + "linear1.weight"].T + sd[p + "linear1.bias"]), drop, base + D_FF_ACT)
        x = x + _drop(a @ sd[p + "linear2.weight"].T + sd[p + "linear2.bias"], drop, base + D_FF_OUT)
    return layer_norm(x, sd["layer_norm.weight"], sd["layer_norm.bias"])


def decoder_forward(sd: Dict[str, Tensor], trg: Tensor, hidden: List[Tensor], n_dec: int, pad_idx: int, n_head: int = 8, drop=None,
                    exits: Optional[List[int]] = None) -> Tensor:
    """early_exit.py:772-800, decoder half: trg (B, L) int64, hidden[e] (B, T', D) -> logits (len(exits), B, L, V)."""
    B, Ln = trg.shape
    x = sd["emb.weight"][trg] + sd["positional_encoder_2.pe"][:Ln, 0, :].to(sd["emb.weight"].dtype)
    x = _drop(x, drop, SITE_PE)
    causal = torch.tril(torch.ones(Ln, Ln, dtype=torch.bool))
    mask = causal[None] & (trg != pad_idx)[:, None, :]              # (B, L, L): key t' visible to query t
    outs = []
    for i, e in enumerate(exits if exits is not None else range(len(hidden))):
        y = decoder_stack(sd, e, n_dec, x, hidden[i if exits is not None else e], mask, n_head, drop)
        outs.append(y @ sd[f"linears_2.{e}.weight"].T + sd[f"linears_2.{e}.bias"])
    return torch.stack(outs, 0)


def aed_loss(dec_out: Tensor, enc_out: Tensor, targets: Tensor, target_lengths: Tensor, ce_weight: float = 0.7, ctc_weight: float = 0.3):
    """train.py:36-51: sum over exits of CE(dec, trg[:, 1:]) (mean over all positions, pad scored) and CTC(enc, full targets)."""
    trg_expect = targets[:, 1:]
    in_len = torch.full((enc_out.shape[1],), enc_out.shape[2], dtype=torch.long)
    ce = sum(torch.nn.functional.cross_entropy(d.permute(0, 2, 1), trg_expect) for d in dec_out)
    ctc = sum(torch.nn.functional.ctc_loss(e.permute(1, 0, 2), targets, in_len, target_lengths, blank=0, zero_infinity=True) for e in enc_out)
    return ce_weight * ce + ctc_weight * ctc, ce, ctc
