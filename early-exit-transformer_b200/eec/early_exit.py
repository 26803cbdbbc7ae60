"""Host-side mirror of the reference's ``models/model/early_exit.py`` for the CTC hot path.

``Early_conformer`` and ``Splitformer`` keep the reference's constructor signature
(early_exit.py:567-582 / :229-244), ``forward(src, lengths) -> (n_exits, B, T', vocab)`` log-probs
(:617-634 / :299-364), ``.train()/.eval()`` semantics and -- by instantiating the same torch.nn
container modules in the same order -- an identical ``state_dict`` (413 / 479 entries) and identical
default initialisation under a given ``torch.manual_seed``.  The containers are never *called*:
``forward`` hands their parameters to the hand-written sm_100a kernels in libeec.so through
``eec.engine``.  There is no PyTorch/CPU fallback.

Precision: ``model.precision = "fp32"`` (default; FFMA kernels, reference-accurate, greedy tokens
bit-exact) or ``"bf16"`` (tcgen05 tensor-core kernels).  ``EEC_PRECISION`` sets the default.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List

import torch
from torch import nn

from . import engine, ops
from .lib import EecError, on_device

Tensor = torch.Tensor


# ------------------------------------------------------------------ parameter containers
class Conv1dSubampling(nn.Module):
    """early_exit.py:24-48 (two Conv1d k=3 s=2, no activation)."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.sequential = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size=3, stride=2, padding=0, padding_mode="zeros"),
            nn.Conv1d(out_channels, out_channels, kernel_size=3, stride=2, padding=0, padding_mode="zeros"),
        )


class PositionalEncoding(nn.Module):
    """models/embedding/positional_encoding.py:55-73."""

    def __init__(self, d_model, dropout, max_len):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class _ConvolutionModule(nn.Module):
    """TA:18-88 parameter layout."""

    def __init__(self, input_dim, num_channels, depthwise_kernel_size, dropout=0.0, bias=False):
        super().__init__()
        if (depthwise_kernel_size - 1) % 2 != 0:
            raise ValueError("depthwise_kernel_size must be odd to achieve 'SAME' padding.")
        self.layer_norm = nn.LayerNorm(input_dim)
        self.sequential = nn.Sequential(
            nn.Conv1d(input_dim, 2 * num_channels, 1, stride=1, padding=0, bias=bias),
            nn.GLU(dim=1),
            nn.Conv1d(num_channels, num_channels, depthwise_kernel_size, stride=1,
                      padding=(depthwise_kernel_size - 1) // 2, groups=num_channels, bias=bias),
            nn.BatchNorm1d(num_channels),
            nn.SiLU(),
            nn.Conv1d(num_channels, input_dim, kernel_size=1, stride=1, padding=0, bias=bias),
            nn.Dropout(dropout),
        )


class _FeedForwardModule(nn.Module):
    """TA:91-119 parameter layout."""

    def __init__(self, input_dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.sequential = nn.Sequential(
            nn.LayerNorm(input_dim), nn.Linear(input_dim, hidden_dim, bias=True), nn.SiLU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim, input_dim, bias=True), nn.Dropout(dropout),
        )


class ConformerLayer(nn.Module):
    """TA:122-174 parameter layout (construction order matters for RNG parity)."""

    def __init__(self, input_dim, ffn_dim, num_attention_heads, depthwise_conv_kernel_size, dropout=0.0):
        super().__init__()
        self.ffn1 = _FeedForwardModule(input_dim, ffn_dim, dropout=dropout)
        self.self_attn_layer_norm = nn.LayerNorm(input_dim)
        self.self_attn = nn.MultiheadAttention(input_dim, num_attention_heads, dropout=dropout)
        self.self_attn_dropout = nn.Dropout(dropout)
        self.conv_module = _ConvolutionModule(input_dim=input_dim, num_channels=input_dim,
                                              depthwise_kernel_size=depthwise_conv_kernel_size, dropout=dropout, bias=True)
        self.ffn2 = _FeedForwardModule(input_dim, ffn_dim, dropout=dropout)
        self.final_layer_norm = nn.LayerNorm(input_dim)


class Conformer(nn.Module):
    """TA:215-271 parameter layout."""

    def __init__(self, input_dim, num_heads, ffn_dim, num_layers, depthwise_conv_kernel_size, dropout=0.0):
        super().__init__()
        self.conformer_layers = nn.ModuleList(
            [ConformerLayer(input_dim, ffn_dim, num_heads, depthwise_conv_kernel_size, dropout=dropout) for _ in range(num_layers)]
        )


# ------------------------------------------------------------------ autograd bridge
class _EncoderFn(torch.autograd.Function):
    """forward(module, want_tape, cfg, want_hidden, src, lengths, *params) -> out | (out, hidden)."""

    @staticmethod
    def forward(ctx, module, want_tape, cfg, want_hidden, src, lengths, *params):
        P = module._tensor_dict()
        side = {"want_hidden": True} if want_hidden else None
        drop_state = None
        if module.training and cfg.drop_p > 0.0:
            # this forward's own copy of {seed, offset}: backward regenerates the masks from it even if another forward
            # ran in between; the module's counter moves on (on device: a graph replay draws new masks every step)
            st = module._dropout_state(src.device)
            drop_state = st.clone()
            with on_device(src.device):
                ops.dropout_advance(st)
        with on_device(src.device):     # kernels launch on the current stream of the TENSORS' device (no global set_device needed)
            out, tape = engine.model_forward(P, module._operands, cfg, src, lengths, module.training, want_tape, side, drop_state)
        ctx.module, ctx.tape, ctx.P, ctx.cfg, ctx.want_hidden = module, tape, P, cfg, want_hidden
        return (out, side["hidden"]) if want_hidden else out

    @staticmethod
    def backward(ctx, gout, ghid=None):
        m = ctx.module
        if ctx.tape is None:
            raise NotImplementedError("eec: backward is only supported in train() mode with grad enabled "
                                      "(eval-mode BatchNorm backward is not implemented)")
        names = m._param_names
        red = m.__dict__.get("_grad_reducer")     # eec.distributed.OverlappedGradReducer: per-exit-group all-reduce during backward
        ctx.cfg._dp_kick = red.kick if (red is not None and getattr(red, "defer", False)) else None
        with on_device(gout.device):
            G = engine.model_backward(ctx.P, m._operands, ctx.cfg, ctx.tape, gout, names,
                                      ghid.contiguous() if ctx.want_hidden and ghid is not None else None,
                                      on_ready=red.on_ready if red is not None else None)
            if red is not None:
                red.finish()
        ctx.tape = None
        m.__dict__["_flat_grad"] = G["__flat__"]  # p.grad tensors are views of this buffer (DP all-reduces it once)
        return (None, None, None, None, None, None) + tuple(G[n] for n in names)


class _EarlyExitBase(nn.Module):
    _splitformer = False

    def __init__(self, src_pad_idx, n_enc_exits, enc_voc_size, dec_voc_size, d_model, n_head, max_len, d_feed_forward,
                 n_enc_layers, features_length, drop_prob, depthwise_kernel_size, device):
        super().__init__()
        self.input_dim = d_model
        self.num_heads = n_head
        self.ffn_dim = d_feed_forward
        self.num_layers = n_enc_layers
        self.depthwise_conv_kernel_size = depthwise_kernel_size
        self.n_enc_exits = n_enc_exits
        self.dropout = drop_prob
        self.device = device
        self.src_pad_idx = src_pad_idx
        self.precision = os.environ.get("EEC_PRECISION", "fp32")
        if d_model % n_head != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.conv_subsample = Conv1dSubampling(in_channels=features_length, out_channels=d_model)
        self.positional_encoder = PositionalEncoding(d_model=d_model, dropout=drop_prob, max_len=max_len)
        self.linears = nn.ModuleList([nn.Linear(d_model, dec_voc_size) for _ in range(n_enc_exits)])
        self.conformer = nn.ModuleList(
            [Conformer(d_model, n_head, d_feed_forward, n_enc_layers, depthwise_kernel_size, dropout=drop_prob)
             for _ in range(n_enc_exits)]
        )
        self._dec_voc_size = dec_voc_size
        self._features_length = features_length
        self.__dict__["_operands_obj"] = None

    # -- kernel specialisation: the hand-written kernels are built for the BASELINE model shape
    def _check_supported(self):
        bad = []
        if self.input_dim != 256: bad.append(f"d_model={self.input_dim} (need 256)")
        if self.num_heads != 8: bad.append(f"n_head={self.num_heads} (need 8)")
        if self.ffn_dim != 2048: bad.append(f"d_feed_forward={self.ffn_dim} (need 2048)")
        if self.depthwise_conv_kernel_size != 31: bad.append(f"depthwise_kernel_size={self.depthwise_conv_kernel_size} (need 31)")
        if self._dec_voc_size != 256: bad.append(f"dec_voc_size={self._dec_voc_size} (need 256)")
        if 3 * self._features_length > 256: bad.append(f"features_length={self._features_length} (need <= 85)")
        if bad:
            raise EecError("eec kernels are specialised for the BASELINE early_conformer shape; unsupported: " + ", ".join(bad))
        if not (0.0 <= float(self.dropout) < 1.0):
            raise ValueError(f"dropout probability has to be between 0 and 1, but got {self.dropout}")
        if self.precision not in ("fp32", "bf16"):
            raise EecError(f"precision must be 'fp32' or 'bf16', got {self.precision!r}")

    def _cfg(self) -> engine.Config:
        bs = self.__dict__.get("_bn_sync")      # eec.distributed.sync_batchnorm(model)
        return engine.Config(n_exits=self.n_enc_exits, n_layers=self.num_layers, n_mels=self._features_length,
                             splitformer=self._splitformer, precision=self.precision, drop_p=float(self.dropout),
                             bn_sync=bs[0] if bs else None, bn_world=bs[1] if bs else 1)

    # -- dropout (train mode, drop_prob > 0): counter-based masks, include/eec.h "dropout".  PyTorch's RNG streams cannot be
    #    matched; the seed follows torch's global seed at first use, so torch.manual_seed(s) before training
    #    makes runs repeatable, and set_dropout_seed() pins it explicitly.
    def _dropout_state(self, dev) -> Tensor:
        st = self.__dict__.get("_drop_state")
        dev = torch.device(dev)
        if st is None or st.device.type != dev.type or (dev.index is not None and st.device.index != dev.index):
            seed = self.__dict__.get("_drop_seed")
            if seed is None:
                seed = torch.initial_seed()
                # data parallel: every rank usually shares torch's seed; give each rank its own mask stream
                if torch.distributed.is_available() and torch.distributed.is_initialized():
                    seed = (seed + torch.distributed.get_rank() * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF
            st = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
            self.__dict__["_drop_state"] = st
        return st

    def set_dropout_seed(self, seed: int, offset: int = 0) -> None:
        """Fix the dropout generator: masks are a pure function of (seed, offset, site, element index); offset counts forwards."""
        self.__dict__["_drop_seed"] = int(seed)
        st = self.__dict__.get("_drop_state")
        if st is not None:
            st.copy_(torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, int(offset)], dtype=torch.int64))

    @property
    def _operands(self) -> engine.Operands:
        ob = self.__dict__["_operands_obj"]
        if ob is None or ob.cfg.precision != self.precision:
            ob = engine.Operands(self._cfg())
            self.__dict__["_operands_obj"] = ob
        return ob

    @property
    def _param_names(self) -> List[str]:
        return [n for n, _ in self.named_parameters()]

    def _tensor_dict(self) -> Dict[str, Tensor]:
        d = {n: p for n, p in self.named_parameters()}
        d.update({n: b for n, b in self.named_buffers()})
        return d

    def forward(self, src: Tensor, lengths: Tensor) -> Tensor:
        self._check_supported()
        params = [p for _, p in self.named_parameters()]
        want_tape = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _EncoderFn.apply(self, want_tape, self._cfg(), False, src, lengths, *params)

    # ---- north-star extension (not in the reference) -------------------------
    @torch.no_grad()
    def forward_early_exit(self, src: Tensor, lengths: Tensor, threshold: float):
        """Dynamic early exit, all decisions on device (no host sync until the caller reads the results).

        After every exit the head kernel emits per-frame argmax and entropy; ``eec_exit_select`` finalises
        utterances whose mean frame entropy over t < length[b] is below ``threshold`` (or all, at the last
        exit), writes their greedy tokens, and compacts the survivors' (T', d) slabs (T' unchanged).
        Returns (exit_index [B] int32, tokens [B,T'] int32 (-1 padded), n_tokens [B] int32, mean_entropy [E,B]).
        """
        from . import early_exit_infer
        self._check_supported()
        if self.training:
            raise EecError("forward_early_exit is an inference API: call model.eval() first")
        with on_device(src.device):
            return early_exit_infer.run(self, src, lengths, threshold)


class Early_conformer(_EarlyExitBase):
    """Drop-in for models.model.early_exit.Early_conformer (early_exit.py:565-634)."""
    _splitformer = False


class Splitformer(_EarlyExitBase):
    """Drop-in for models.model.early_exit.Splitformer (early_exit.py:227-364)."""
    _splitformer = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.factor = 2
        self.conformer_parallel = nn.ModuleList(
            [Conformer(self.input_dim, self.num_heads, self.ffn_dim, 1, self.depthwise_conv_kernel_size, dropout=self.dropout)
             for _ in range(2)]
        )


def greedy_decode(log_probs: Tensor, blank: int = 0):
    """GreedyCTCDecoder (util/beam_infer.py:9-24) for a whole (B,T,V) or (E,B,T,V) tensor of log-probs on
    device: argmax -> collapse repeats -> drop blank, over ALL T frames.  Returns (tokens [...,T] int32
    padded with -1, n_tokens [...] int32)."""
    if not log_probs.is_cuda:
        raise EecError("greedy_decode: tensor must be on a CUDA device")
    shp = log_probs.shape
    lp = log_probs.contiguous().view(-1, shp[-2], shp[-1]).float()
    Bx, T, Vv = lp.shape
    dev = lp.device
    am = torch.empty(Bx * T, dtype=torch.int32, device=dev)
    scratch = torch.empty_like(lp)
    tokens = torch.empty(Bx, T, dtype=torch.int32, device=dev)
    n_tok = torch.empty(Bx, dtype=torch.int32, device=dev)
    with on_device(dev):
        ops.call("eec_logsoftmax_fwd", ops.ptr(lp), ops.ptr(scratch), ops.ptr(am), None, Bx * T, Vv, ops.stream())
        ops.greedy_collapse(am, tokens, n_tok, Bx, T, blank)
    return tokens.view(*shp[:-2], T), n_tok.view(*shp[:-2])
