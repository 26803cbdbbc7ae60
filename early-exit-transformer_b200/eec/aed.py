"""``full_conformer`` (reference early_exit.py:637-811): the AED-mode model of BASELINE configs[4].

SURVEY §8 row a17 puts the ENCODER half on the hot path: front end, positional encoding, the six Conformer groups
and the ``linears_1`` CTC heads are the same arithmetic as ``Early_conformer`` and run on the sm_100a kernels; the
per-exit encoder states additionally leave the engine (fp32, [E,B,T',256]) because the attention decoders consume
them, and their gradient flows back into the engine's backward pass.  The six ``nn.TransformerDecoder`` stacks,
``linears_2`` and the embedding are row N4 ("next"): they stay the same torch.nn library modules the reference
instantiates (early_exit.py:701-717), called exactly as the reference calls them (:742-762, :772-798).

Same constructor signature, ``forward(src, lengths, trg) -> (dec_out (E,B,L,V) logits, enc_out (E,B,T',V) log-probs)``,
``_encoder_(src, lengths, layer_n)`` / ``_decoder_(trg, enc, layer_n)`` (used by the reference's AED beam search,
inference.py:18-62), module construction order (identical default init under a seed) and state_dict layout.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch
from torch import nn

from . import engine
from .early_exit import Conformer, Conv1dSubampling, PositionalEncoding, _EarlyExitBase, _EncoderFn

Tensor = torch.Tensor

# engine name (Early_conformer layout) <- full_conformer state_dict name
_RENAME = (("linears_1.", "linears."), ("positional_encoder_1.", "positional_encoder."))
_ENCODER_PREFIXES = ("conv_subsample.", "linears_1.", "positional_encoder_1.", "conformer.")


def _engine_name(name: str) -> str:
    for a, b in _RENAME:
        if name.startswith(a):
            return b + name[len(a):]
    return name


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
        """early_exit.py:739-762 (torch.nn decoder stack `layer_n`, log-softmax output)."""
        dev = enc.device
        tgt_mask = self.create_tgt_mask(trg.size(1)).to(dev)
        tgt_key_padding_mask = self.create_pad_mask(trg, self.trg_pad_idx).to(dev)
        trg = self.emb(trg)
        trg = self._pos2(trg)
        i = (int(layer_n) if 1 <= int(layer_n) <= self.n_enc_exits else self.n_enc_exits) - 1   # (:751-755: same loop quirk)
        out_d = self.decoders[i](trg, enc, tgt_mask=tgt_mask, tgt_key_padding_mask=tgt_key_padding_mask)
        return torch.nn.functional.log_softmax(self.linears_2[i](out_d), dim=2)

    def _pos2(self, x: Tensor) -> Tensor:
        # positional_encoding.py:65-73 on a batch-first (B, L, D) tensor: permute -> + pe[:L] -> permute back -> dropout
        pe = self.positional_encoder_2
        return pe.dropout(x + pe.pe[: x.size(1), 0, :].unsqueeze(0))

    def forward(self, src: Tensor, lengths: Tensor, trg: Tensor):
        """early_exit.py:764-800 -> (dec_out [E,B,L,V] logits, enc_out [E,B,T',V] log-probs)."""
        enc_out, hidden = self._encode(src, lengths, self.n_enc_exits)
        dev = hidden.device
        tgt_mask = self.create_tgt_mask(trg.size(1)).to(dev)
        tgt_key_padding_mask = self.create_pad_mask(trg, self.trg_pad_idx).to(dev)
        t = self._pos2(self.emb(trg))
        dec_out = []
        for e, (linear_2, decoder) in enumerate(zip(self.linears_2, self.decoders)):
            out_d = decoder(t, hidden[e], tgt_mask=tgt_mask, tgt_key_padding_mask=tgt_key_padding_mask)
            dec_out.append(linear_2(out_d).unsqueeze(0))
        return torch.cat(dec_out), enc_out

    def create_pad_mask(self, matrix: Tensor, pad_token: int) -> Tensor:
        return matrix == pad_token

    def create_tgt_mask(self, sz: int) -> Tensor:
        return torch.triu(torch.full((sz, sz), float("-inf")), diagonal=1)

    def forward_early_exit(self, *a, **k):
        raise NotImplementedError("forward_early_exit is a CTC-mode API (Early_conformer / Splitformer)")
