"""``full_conformer`` (reference early_exit.py:637-811): the AED-mode model of BASELINE configs[4].

Encoder half: front end, positional encoding, the six Conformer groups and the ``linears_1`` CTC heads are the
same arithmetic as ``Early_conformer`` and run on the sm_100a kernels through ``eec.engine``; the per-exit encoder states leave the
engine (fp32, [E,B,T',256]) for the decoders, and their gradient flows back into the engine's backward pass.
Decoder half (row N4): embedding + positional encoding, the six pre-norm ``nn.TransformerDecoder`` stacks (causal self-attention with
the target padding mask, cross-attention over the encoder states, ReLU feed-forward, ONE shared final LayerNorm) and ``linears_2`` run
on the same kernels through ``eec.decoder_engine``.  The torch.nn modules below are parameter containers only (same construction order
as early_exit.py:667-717, so default init under a seed and the state_dict layout are the reference's); they are never called.

Same constructor signature, ``forward(src, lengths, trg) -> (dec_out (E,B,L,V) logits, enc_out (E,B,T',V) log-probs)``,
``_encoder_(src, lengths, layer_n)`` / ``_decoder_(trg, enc, layer_n)`` (used by the reference's AED beam search,
inference.py:18-62), module construction order (identical default init under a seed) and state_dict layout.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch
from torch import nn

from . import decoder_engine, engine, ops
from .early_exit import Conformer, Conv1dSubampling, PositionalEncoding, _EarlyExitBase, _EncoderFn
from .lib import EecError, on_device

Tensor = torch.Tensor

# engine name (Early_conformer layout) <- full_conformer state_dict name
_RENAME = (("linears_1.", "linears."), ("positional_encoder_1.", "positional_encoder."))
_ENCODER_PREFIXES = ("conv_subsample.", "linears_1.", "positional_encoder_1.", "conformer.")


def _engine_name(name: str) -> str:
    for a, b in _RENAME:
        if name.startswith(a):
            return b + name[len(a):]
    return name


class _DecoderFn(torch.autograd.Function):
    """forward(module, want_tape, cfg, exits, mem_index, trg, hidden, *decoder params) -> logits [len(exits), B, L, V]"""

    @staticmethod
    def forward(ctx, module, want_tape, cfg, exits, mem_index, trg, hidden, *params):
        P = module._decoder_tensor_dict()
        drop0 = None
        if module.training and cfg.drop_p > 0.0:
            # this forward's own {seed, offset} (the decoder shares the module's counter with the encoder half: every forward, encoder or
            # decoder, advances it, and backward regenerates its masks from the snapshot)
            st = module._dropout_state(hidden.device)
            with on_device(hidden.device):
                drop0 = ops.Drop(st.clone(), cfg.drop_p, 0)
                ops.dropout_advance(st)
        with on_device(hidden.device):
            out, tape = decoder_engine.decoder_forward(P, module._dec_operands, cfg, module.n_dec_layers, trg.to(hidden.device),
                                                       module.trg_pad_idx, hidden.contiguous(), exits, mem_index, want_tape, drop0)
        ctx.module, ctx.tape, ctx.P, ctx.cfg, ctx.hshape = module, tape, P, cfg, tuple(hidden.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        m = ctx.module
        if ctx.tape is None:
            raise NotImplementedError("eec: decoder backward needs train() mode with grad enabled")
        names = m._decoder_param_names
        ghid = torch.zeros(ctx.hshape, dtype=torch.float32, device=gout.device)
        with on_device(gout.device):
            G = decoder_engine.decoder_backward(ctx.P, m._dec_operands, ctx.cfg, ctx.tape, gout, names, ghid)
        ctx.tape = None
        m.__dict__["_flat_grad_decoder"] = G["__flat__"]
        return (None, None, None, None, None, None, ghid) + tuple(G[n] for n in names)


class _CeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_eblv, targets):
        E, B, Ln, Vv = logits_eblv.shape
        x = logits_eblv.contiguous()
        tg = targets.to(device=x.device, dtype=torch.int64).contiguous().view(-1)
        loss = torch.zeros(E, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        with on_device(x.device):
            for e in range(E):
                ops.cross_entropy(x[e].view(B * Ln, Vv), tg, loss[e:e + 1], grad[e].view(B * Ln, Vv) if grad is not None else None)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, gloss):
        g = ctx.grad
        with on_device(g.device):
            ops.scale_rows_dev(g, gloss.contiguous().float(), g)
        ctx.grad = None
        return g, None


def multi_exit_cross_entropy(dec_out: Tensor, trg_expect: Tensor, reduce_exits: bool = True) -> Tensor:
    """train.py:47 for all exits: sum_e nn.CrossEntropyLoss()(dec_out[e].permute(0, 2, 1), trg_expect) -- mean over ALL B*L positions of every
    exit, no ignore_index (the pad id is scored); one fused log-softmax + NLL (+ gradient) kernel per exit."""
    if not dec_out.is_cuda or dec_out.dim() != 4 or dec_out.dtype != torch.float32:
        raise EecError("multi_exit_cross_entropy: expected (E, B, L, V) fp32 logits on a CUDA device")
    per_exit = _CeFn.apply(dec_out, trg_expect)
    return per_exit.sum() if reduce_exits else per_exit


class full_conformer(_EarlyExitBase):
    """Drop-in for models.model.early_exit.full_conformer (early_exit.py:637-811)."""
    _splitformer = False

    def __init__(self, trg_pad_idx, n_enc_exits, enc_voc_size, dec_voc_size, d_model, n_head, max_len, d_feed_forward,
                 n_enc_layers, n_dec_layers, features_length, drop_prob, depthwise_kernel_size, device):
        nn.Module.__init__(self)
        self.input_dim = d_model
        self.num_heads = n_head
        self.ffn_dim = d_feed_forward
        self.num_layers = n_enc_layers
        self.depthwise_conv_kernel_size = depthwise_kernel_size
        self.n_enc_exits = n_enc_exits
        self.dropout = drop_prob
        self.n_dec_layers = n_dec_layers
        self.device = device
        self.precision = os.environ.get("EEC_PRECISION", "fp32")
        if d_model % n_head != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        # construction order = the reference's (early_exit.py:667-717): same RNG consumption, same state_dict order
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-5)
        self.emb = nn.Embedding(dec_voc_size, d_model)
        self.trg_pad_idx = trg_pad_idx
        self.conv_subsample = Conv1dSubampling(in_channels=features_length, out_channels=d_model)
        self.linears_1 = nn.ModuleList([nn.Linear(d_model, dec_voc_size) for _ in range(n_enc_exits)])
        self.linears_2 = nn.ModuleList([nn.Linear(d_model, dec_voc_size) for _ in range(n_enc_exits)])
        self.positional_encoder_1 = PositionalEncoding(d_model=d_model, dropout=drop_prob, max_len=max_len)
        self.positional_encoder_2 = PositionalEncoding(d_model=d_model, dropout=drop_prob, max_len=max_len)
        self.conformer = nn.ModuleList(
            [Conformer(d_model, n_head, d_feed_forward, n_enc_layers, depthwise_kernel_size, dropout=drop_prob)
             for _ in range(n_enc_exits)]
        )
        self.decoders = nn.ModuleList(
            [nn.TransformerDecoder(
                nn.TransformerDecoderLayer(d_model=d_model, nhead=n_head, dim_feedforward=d_feed_forward, dropout=drop_prob,
                                           batch_first="True", norm_first="True"),
                n_dec_layers, self.layer_norm)
             for _ in range(n_enc_exits)]
        )
        self._dec_voc_size = dec_voc_size
        self._features_length = features_length
        self.__dict__["_operands_obj"] = None

    # ---- the encoder half, as the engine sees it (Early_conformer parameter names) -------------
    def _encoder_named_parameters(self):
        return [(n, p) for n, p in self.named_parameters() if n.startswith(_ENCODER_PREFIXES)]

    @property
    def _param_names(self) -> List[str]:
        return [_engine_name(n) for n, _ in self._encoder_named_parameters()]

    def _tensor_dict(self) -> Dict[str, Tensor]:
        d = {_engine_name(n): p for n, p in self._encoder_named_parameters()}
        d.update({_engine_name(n): b for n, b in self.named_buffers() if n.startswith(_ENCODER_PREFIXES)})
        return d

    # ---- the decoder half, as eec.decoder_engine sees it -------------------------------------------------------------------------
    @property
    def _decoder_param_names(self) -> List[str]:
        return [n for n, _ in self.named_parameters() if not n.startswith(_ENCODER_PREFIXES)]

    def _decoder_tensor_dict(self) -> Dict[str, Tensor]:
        d = {n: p for n, p in self.named_parameters() if not n.startswith(_ENCODER_PREFIXES)}
        d["positional_encoder_2.pe"] = self.positional_encoder_2.pe
        return d

    @property
    def _dec_operands(self) -> engine.Operands:
        ob = self.__dict__.get("_dec_operands_obj")
        if ob is None or ob.cfg.precision != self.precision:
            ob = engine.Operands(self._cfg())
            self.__dict__["_dec_operands_obj"] = ob
        return ob

    def _decode(self, trg: Tensor, hidden: Tensor, exits: List[int], mem_index: List[int]) -> Tensor:
        """decoder stacks `exits` (0-based), stack exits[i] attending to hidden[mem_index[i]] -> logits [len(exits), B, L, V]"""
        self._check_supported()
        names = self._decoder_param_names
        expect = decoder_engine.decoder_param_names(self.n_enc_exits, self.n_dec_layers)
        if names != expect:
            raise EecError("eec.full_conformer: unexpected decoder parameter layout")
        P = dict(self.named_parameters())
        params = [P[n] for n in names]
        want_tape = self.training and torch.is_grad_enabled() and (hidden.requires_grad or any(p.requires_grad for p in params))
        return _DecoderFn.apply(self, want_tape, self._cfg(), list(exits), list(mem_index), trg, hidden, *params)

    def _encode(self, src: Tensor, lengths: Tensor, n_exits: int):
        """-> (enc_out [n_exits,B,T',V] log-probs, hidden [n_exits,B,T',D]) for the first `n_exits` groups."""
        self._check_supported()
        params = [p for _, p in self._encoder_named_parameters()]
        want_tape = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        full = self._cfg()
        cfg = engine.Config(n_exits=n_exits, n_layers=full.n_layers, n_mels=full.n_mels, precision=full.precision, drop_p=full.drop_p)
        return _EncoderFn.apply(self, want_tape, cfg, True, src, lengths, *params)

    def _encoder_(self, src: Tensor, lengths: Tensor, layer_n: int) -> Tensor:
        """early_exit.py:719-737: encoder state after exit group `layer_n` (1-based).  The reference's loop only breaks when its counter
        EQUALS layer_n, so layer_n <= 0 or > n_enc_exits runs ALL groups."""
        n = int(layer_n) if 1 <= int(layer_n) <= self.n_enc_exits else self.n_enc_exits
        _, hidden = self._encode(src, lengths, n)
        return hidden[n - 1]

    def _decoder_(self, trg: Tensor, enc: Tensor, layer_n: int) -> Tensor:
        """early_exit.py:739-762: decoder stack `layer_n` (1-based) on the encoder state `enc` (B, T', D), log-softmax output."""
        i = (int(layer_n) if 1 <= int(layer_n) <= self.n_enc_exits else self.n_enc_exits) - 1   # (:751-755: same loop quirk)
        logits = self._decode(trg, enc.unsqueeze(0), [i], [0])[0]
        B, Ln, Vv = logits.shape
        out = torch.empty_like(logits)
        with on_device(logits.device):
            ops.call("eec_logsoftmax_fwd", ops.ptr(logits.contiguous()), ops.ptr(out), None, None, B * Ln, Vv, ops.stream())
        return out

    def forward(self, src: Tensor, lengths: Tensor, trg: Tensor):
        """early_exit.py:764-800 -> (dec_out [E,B,L,V] logits, enc_out [E,B,T',V] log-probs)."""
        enc_out, hidden = self._encode(src, lengths, self.n_enc_exits)
        idx = list(range(self.n_enc_exits))
        return self._decode(trg, hidden, idx, idx), enc_out

    def create_pad_mask(self, matrix: Tensor, pad_token: int) -> Tensor:
        return matrix == pad_token

    def create_tgt_mask(self, sz: int) -> Tensor:
        return torch.triu(torch.full((sz, sz), float("-inf")), diagonal=1)

    def forward_early_exit(self, *a, **k):
        raise NotImplementedError("forward_early_exit is a CTC-mode API (Early_conformer / Splitformer)")
