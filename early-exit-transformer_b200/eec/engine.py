"""Forward / backward orchestration of the early-exit conformer encoder over libeec.so kernels.

This is the "manual tape" engine behind ``eec.early_exit.Early_conformer``: one Python function
launches every kernel of the forward pass (and records the buffers backward needs), a second one
replays the exact adjoint.  No autograd graph, no torch arithmetic: torch tensors are device
buffers only.  Layer arithmetic follows torchaudio ``ConformerLayer.forward`` (TA:176-212) as
restated in SURVEY.md Appendix A; model structure follows early_exit.py:617-634 (Early_conformer)
and :299-364 (Splitformer).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import os
from typing import Dict, List, Optional

import torch

from . import ops
from .lib import ACT_DSILU, ACT_GLU, ACT_NONE, ACT_SILU, EecError

Tensor = torch.Tensor
D, F, H, KW, V = 256, 2048, 8, 31, 256
BN_MOMENTUM = 0.1

MATRIX_SUFFIXES = (
    "ffn1.sequential.1.weight", "ffn1.sequential.4.weight", "ffn2.sequential.1.weight", "ffn2.sequential.4.weight",
    "self_attn.in_proj_weight", "self_attn.out_proj.weight", "conv_module.sequential.0.weight",
    "conv_module.sequential.5.weight",
)


@dataclass
class Config:
    n_exits: int
    n_layers: int          # per exit group
    n_mels: int = 80
    splitformer: bool = False
    precision: str = "fp32"   # "fp32" (FFMA parity path) or "bf16" (tcgen05 path)
    drop_p: float = 0.0       # train-mode dropout probability (the reference's --drop_prob; applied when a drop state is passed)
    bn_sync: Optional[object] = None   # data parallel with synchronised BatchNorm (eec.distributed.sync_batchnorm): callable that
    bn_world: int = 1                  #   SUM-all-reduces a double tensor in place over bn_world ranks of equal batch shape
    total_exits: int = 0      # exits of the FULL model when n_exits is a truncation (Splitformer: the parallel branches sit at group 0 and
                              # at the full model's last group, early_exit.py:314 / :340); 0 = n_exits

    @property
    def last_exit(self) -> int:
        return (self.total_exits or self.n_exits) - 1

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class _ZeroPool:
    """Per-layer BatchNorm statistic accumulators (2 x 256 doubles each) carved out of ONE zero-filled buffer per forward / backward:
    one fill launch instead of one per layer (24 launches of a 12-layer step)."""

    def __init__(self, n: int, dev):
        self.buf = torch.zeros(max(n, 1), 2 * D, dtype=torch.float64, device=dev)
        self.i = 0

    def take(self) -> Tensor:
        if self.i >= self.buf.shape[0]:
            return torch.zeros(2 * D, dtype=torch.float64, device=self.buf.device)
        self.i += 1
        return self.buf[self.i - 1]


class Operands:
    """GEMM-operand views of the parameters in the activation dtype.  fp32: zero-copy reshapes;
    bf16: cast by eec_cast into buffers that are refreshed when a parameter's version changes."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self._cache: Dict[str, tuple] = {}
        self._shadow: Optional[Tensor] = None      # flat bf16 copy of all parameters maintained by eec.FusedNoamAdamW
        self._shadow_off: Dict[str, tuple] = {}
        self._shadow_ver: Dict[str, int] = {}

    def attach_shadow(self, shadow: Tensor, offsets: Dict[str, tuple], versions: Dict[str, int]) -> None:
        self._shadow, self._shadow_off, self._shadow_ver = shadow, offsets, versions

    def invalidate(self) -> None:
        """Parameters changed through a raw pointer (fused optimiser): drop every cached cast."""
        self._cache.clear()

    def get(self, name: str, p: Tensor, shape2d) -> Tensor:
        if self.cfg.precision == "fp32":
            return p.detach().reshape(shape2d)
        if self._shadow is not None and self._shadow_ver.get(name) == p._version and name in self._shadow_off:
            off, k = self._shadow_off[name]
            return self._shadow[off:off + k].view(shape2d)   # kept current by the optimiser's own update pass
        key = (p.data_ptr(), p._version)
        hit = self._cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        buf = hit[1] if hit is not None else _empty(shape2d, torch.bfloat16, p.device)
        ops.cast(p.detach(), buf)
        self._cache[name] = (key, buf)
        return buf

    def front_w1(self, p: Tensor) -> Tensor:
        """conv1 weight (256, n_mels, 3) -> zero-padded (256, 256) operand (K = 3*n_mels padded to 256)."""
        name = "conv_subsample.sequential.0.weight"
        key = (p.data_ptr(), p._version)
        hit = self._cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        k = p.shape[1] * 3
        staged = torch.zeros(256, 256, dtype=torch.float32, device=p.device)
        staged[:, :k].copy_(p.detach().reshape(256, k))  # layout plumbing only
        if self.cfg.precision == "fp32":
            buf = staged
        else:
            buf = hit[1] if hit is not None else _empty((256, 256), torch.bfloat16, p.device)
            ops.cast(staged, buf)
        self._cache[name] = (key, buf)
        return buf


# ----------------------------------------------------------------------------------------------
# small helpers over ops.gemm
# ----------------------------------------------------------------------------------------------
def linear(A, W, out, M, N, K, **kw):
    ops.gemm(A, W, out, M, N, K, a_kmajor=True, b_kmajor=True, **kw)


def dgrad(dY, W, out, M, N_out, K_in, **kw):
    """out[M,K_in] = dY[M,N_out] @ W[N_out,K_in]"""
    ops.gemm(dY, W, out, M, K_in, N_out, a_kmajor=True, b_kmajor=False, lda=N_out, ldb=K_in, **kw)


_BF16_DU = os.environ.get("EEC_BF16_DU", "1") != "0"   # 0: keep the LayerNorm-backward inputs in fp32 (A/B runs)
_FUSE_COLSUM = not (os.environ.get("EEC_GEMM_V1") == "1" or os.environ.get("EEC_GEMM_V2") == "1" or os.environ.get("EEC_FORCE_SIMT") == "1")


class SideQueue:
    """Weight-gradient GEMMs off the critical path.  The data-gradient chain of backward (dgrad -> LayerNorm backward -> dgrad ...)
    never reads a weight gradient, so every wgrad is issued on a second stream: it starts as soon as its dY exists (event fork)
    and only has to finish before the gradients are consumed (join before the data-parallel hook of its exit group / at the end of
    backward).  Inside the step's CUDA graph this is a parallel branch: the wgrad CTAs fill the tails and launch gaps of the
    chain's kernels and co-reside with its bandwidth-bound ones.  Operand tensors are kept referenced until the join, so the
    caching allocator cannot hand their memory to a later allocation of the compute stream while the side stream still reads them.
    bf16 path only (there every dY operand is a private copy; the fp32 path feeds the in-place residual gradient)."""

    def __init__(self, dev):
        side = _SIDE_STREAMS.get(("wgrad", dev))
        if side is None:
            side = _SIDE_STREAMS[("wgrad", dev)] = torch.cuda.Stream(device=dev)
        self.dev, self.side, self.refs = dev, side, []

    def run(self, fn, *tensors):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()
        self.refs.extend(tensors)

    def join(self):
        torch.cuda.current_stream(self.dev).wait_stream(self.side)
        self.refs.clear()


_WGRAD_SIDE = os.environ.get("EEC_WGRAD_SIDE", "1") != "0"


def wgrad(dY, X, dW, M, N_out, K_in, alpha=1.0, dbias=None, sq: Optional["SideQueue"] = None):
    """dW[N_out,K_in] += alpha * dY[M,N_out]^T @ X[M,K_in];  dbias[N_out] += column sums of dY (optional).
    On the tcgen05 path the bias gradient is summed from the dY tiles while they sit in shared memory as the GEMM's
    A operand (eec_gemm_desc.a_colsum): the [M, N_out] tensor is not read a second time.
    sq: run on the side stream (SideQueue)."""
    fuse = dbias is not None and _FUSE_COLSUM and dY.dtype == torch.bfloat16

    def go():
        ops.gemm(dY, X, dW, N_out, K_in, M, a_kmajor=False, b_kmajor=False, lda=N_out, ldb=K_in, alpha=alpha, accumulate=True,
                 a_colsum=dbias if fuse else None)
        if dbias is not None and not fuse:
            ops.colsum(dY, dbias, M, N_out)
    if sq is not None and dY.dtype == torch.bfloat16:
        sq.run(go, dY, X)
    else:
        go()


def to_act(x32: Tensor, cfg: Config) -> Tensor:
    if cfg.precision == "fp32":
        return x32
    out = _empty(x32.shape, torch.bfloat16, x32.device)
    ops.cast(x32, out)
    return out


# ----------------------------------------------------------------------------------------------
# one ConformerLayer
# ----------------------------------------------------------------------------------------------
# dropout sites of one ConformerLayer, as offsets from the layer's site base (ops.Drop.at): the nn.Dropout modules at
# TA:106 / TA:108 (both FFNs), nn.MultiheadAttention's dropout on the probabilities (TA:152), TA:201, TA:73
S_FFN1_ACT, S_FFN1_OUT, S_ATTN_P, S_ATTN_OUT, S_CONV_OUT, S_FFN2_ACT, S_FFN2_OUT = range(7)
SITE_PE = 0                      # positional_encoding.py:72
SITES_PER_LAYER = 8


def _at(drop, k):
    return drop.at(k) if drop is not None else None


def site_geometry(B: int, T: int):
    """logical tensor (R, C, Cs, W) of every dropout site of a layer run on B utterances of T frames (include/eec.h eec_dropout_bits)"""
    N = B * T
    g = {k: (N, D, D, 32) for k in (S_FFN1_OUT, S_ATTN_OUT, S_CONV_OUT, S_FFN2_OUT)}
    g[S_FFN1_ACT] = g[S_FFN2_ACT] = (N, F, F, 16)
    g[S_ATTN_P] = (B * H * T, T, 8 * ((T + 7) // 8), 32)
    return g


_SIDE_STREAMS: Dict[torch.device, "torch.cuda.Stream"] = {}


class MaskPlan:
    """Keep-mask words of every dropout site of one forward (tensor-core path), generated on a SIDE stream: the generator is
    integer-ALU work that depends on nothing but the forward's {seed, counter}, while the compute stream runs tensor-core / TMA
    bound kernels -- so all layers' words are issued up front on a second stream (a parallel branch of the step's CUDA graph)
    and each layer waits for its own event.  Buffers are allocated on the compute stream first: the side stream only ever
    touches them before the event the compute stream waits on, so the caching allocator's single-stream bookkeeping stays valid."""

    def __init__(self, drop0: ops.Drop, layers):   # layers: [(uid, B, T)] in execution order
        dev = drop0.state.device
        cur = torch.cuda.current_stream(dev)
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
        self.sites, self.events, fills = {}, {}, {}
        for uid, B, T in layers:
            base = drop0.at((1 + uid) * SITES_PER_LAYER)
            self.sites[uid], fills[uid] = {}, []
            for k, geo in site_geometry(B, T).items():
                d, fill = base.at(k).alloc_bits(*geo)
                self.sites[uid][k] = d
                fills[uid].append(fill)
        side.wait_stream(cur)          # (the drop state was written on the compute stream)
        with torch.cuda.stream(side):
            for uid, _, _ in layers:
                for fill in fills[uid]:
                    fill()
                ev = torch.cuda.Event()
                ev.record(side)
                self.events[uid] = ev
        self.side = side

    def take(self, uid: int):
        """-> {site: Drop with words}; the compute stream waits for this layer's words"""
        torch.cuda.current_stream(self.side.device).wait_event(self.events[uid])
        return self.sites[uid]

    def join(self):
        torch.cuda.current_stream(self.side.device).wait_stream(self.side)


def layer_forward(P: Dict[str, Tensor], W: Operands, pre: str, x: Tensor, key_len: Tensor, B: int, T: int, cfg: Config,
                  training: bool, tape: Optional[dict], drop: Optional[ops.Drop] = None, masks: Optional[dict] = None):
    """x: fp32 [B*T, 256] -> fp32 [B*T, 256].  TA:176-212.  `drop` = this layer's dropout site base (train mode, p > 0)."""
    N = B * T
    dev = x.device
    TD = cfg.act_dtype
    f32 = torch.float32
    save = tape is not None

    def stat():
        return (_empty((N,), f32, dev), _empty((N,), f32, dev)) if save else (None, None)

    bf16 = cfg.precision == "bf16"
    dsites = {}      # dropout sites whose keep-mask words the backward kernels read again (tensor-core path)

    def site(k: int, R: int = 0, C: int = 0, Cs: int = 0, Wd: int = 0, keep: bool = False):
        d = _at(drop, k)
        if d is not None and bf16 and Wd:
            d = masks[k] if masks is not None else d.with_bits(R, C, Cs, Wd)   # (MaskPlan: generated ahead on the side stream)
            if keep:
                dsites[k] = d
        return d

    def ffn(tag: str, x_in: Tensor, u: Tensor, ln_next_g, ln_next_b, ln_out_dtype, s_act: int, s_out: int):
        q = pre + tag + ".sequential."
        W1 = W.get(q + "1.weight", P[q + "1.weight"], (F, D))
        W2 = W.get(q + "4.weight", P[q + "4.weight"], (D, F))
        hpre = _empty((N, F), TD, dev) if save else None
        a = _empty((N, F), TD, dev)
        linear(u, W1, a, N, F, D, bias=P[q + "1.bias"], act=ACT_SILU, preact=hpre, drop=site(s_act, N, F, F, 16, keep=save))
        x_out = _empty((N, D), f32, dev)
        u_next = _empty((N, D), ln_out_dtype, dev)
        m, r = stat()
        linear(a, W2, x_out, N, D, F, bias=P[q + "4.bias"], alpha=0.5, residual=x_in, ln_gamma=ln_next_g, ln_beta=ln_next_b,
               ln_out=u_next, ln_mean=m, ln_rstd=r, drop=site(s_out, N, D, D, 32))
        return x_out, u_next, hpre, a, m, r

    # ---- FFN1 (TA:185-187): x1 = x + 0.5*FFN(LN(x));  u2 = LN_attn(x1) fused into the GEMM tail
    q = pre + "ffn1.sequential."
    u1 = _empty((N, D), TD, dev)
    m1, r1 = stat()
    ops.layernorm_fwd(x, P[q + "0.weight"], P[q + "0.bias"], u1, m1, r1)
    x1, u2, h1pre, a1, m2, r2 = ffn("ffn1", x, u1, P[pre + "self_attn_layer_norm.weight"], P[pre + "self_attn_layer_norm.bias"], TD,
                                    S_FFN1_ACT, S_FFN1_OUT)

    # ---- MHSA (TA:192-202)
    Wqkv = W.get(pre + "self_attn.in_proj_weight", P[pre + "self_attn.in_proj_weight"], (3 * D, D))
    Wo = W.get(pre + "self_attn.out_proj.weight", P[pre + "self_attn.out_proj.weight"], (D, D))
    qkv = _empty((N, 3 * D), TD, dev)
    linear(u2, Wqkv, qkv, N, 3 * D, D, bias=P[pre + "self_attn.in_proj_bias"])
    ctx = _empty((N, D), TD, dev)
    lse = _empty((B, H, T), f32, dev)
    ops.attn_fwd(qkv, key_len, ctx, lse, B, T, H, drop=site(S_ATTN_P, B * H * T, T, 8 * ((T + 7) // 8), 32, keep=save))
    x2 = _empty((N, D), f32, dev)
    u3 = _empty((N, D), TD, dev)
    m3, r3 = stat()
    c = pre + "conv_module."
    linear(ctx, Wo, x2, N, D, D, bias=P[pre + "self_attn.out_proj.bias"], residual=x1, ln_gamma=P[c + "layer_norm.weight"],
           ln_beta=P[c + "layer_norm.bias"], ln_out=u3, ln_mean=m3, ln_rstd=r3, drop=site(S_ATTN_OUT, N, D, D, 32))

    # ---- convolution module (TA:42-75, 168-174)
    Wp1 = W.get(c + "sequential.0.weight", P[c + "sequential.0.weight"], (2 * D, D))
    Wp2 = W.get(c + "sequential.5.weight", P[c + "sequential.5.weight"], (D, D))
    z = _empty((N, 2 * D), TD, dev) if (save or cfg.precision == "fp32") else None
    g = _empty((N, D), TD, dev)
    linear(u3, Wp1, g, N, 2 * D, D, bias=P[c + "sequential.0.bias"], act=ACT_GLU, preact=z)
    s = _empty((N, D), TD, dev)
    wdw = P[c + "sequential.2.weight"].detach().reshape(D, KW)
    cbuf = sm = sr = None
    if training:
        cbuf = _empty((N, D), f32, dev)
        zp = getattr(cfg, "_zero_pool", None)
        sums = zp.take() if zp is not None else torch.zeros(2 * D, dtype=torch.float64, device=dev)
        ops.dwconv_stats(g, wdw, P[c + "sequential.2.bias"], cbuf, sums, B, T, KW)
        if cfg.bn_sync is not None:
            cfg.bn_sync(sums)          # per-channel sum / sum of squares over the GLOBAL batch (TA:59 BatchNorm1d under sync-BN)
        sm, sr = _empty((D,), f32, dev), _empty((D,), f32, dev)
        ops.bn_silu_train(cbuf, sums, P[c + "sequential.3.weight"], P[c + "sequential.3.bias"],
                          P[c + "sequential.3.running_mean"], P[c + "sequential.3.running_var"],
                          P[c + "sequential.3.num_batches_tracked"], BN_MOMENTUM, sm, sr, s,
                          stat_rows=N * cfg.bn_world if cfg.bn_sync is not None else 0)
    else:
        ops.dwconv_bn_silu_eval(g, wdw, P[c + "sequential.2.bias"], P[c + "sequential.3.weight"], P[c + "sequential.3.bias"],
                                P[c + "sequential.3.running_mean"], P[c + "sequential.3.running_var"], s, B, T, KW)
    x3 = _empty((N, D), f32, dev)
    u4 = _empty((N, D), TD, dev)
    m4, r4 = stat()
    q2 = pre + "ffn2.sequential."
    linear(s, Wp2, x3, N, D, D, bias=P[c + "sequential.5.bias"], residual=x2, ln_gamma=P[q2 + "0.weight"],
           ln_beta=P[q2 + "0.bias"], ln_out=u4, ln_mean=m4, ln_rstd=r4, drop=site(S_CONV_OUT, N, D, D, 32))

    # ---- FFN2 + final LayerNorm (TA:207-211): x4 = x3 + 0.5*FFN(u4); y = LN_final(x4)
    x4, y, h2pre, a2, m5, r5 = ffn("ffn2", x3, u4, P[pre + "final_layer_norm.weight"], P[pre + "final_layer_norm.bias"], f32,
                                   S_FFN2_ACT, S_FFN2_OUT)

    if save:
        tape.update(dict(x=x, u1=u1, m1=m1, r1=r1, h1pre=h1pre, a1=a1, x1=x1, u2=u2, m2=m2, r2=r2, qkv=qkv, ctx=ctx, lse=lse,
                         x2=x2, u3=u3, m3=m3, r3=r3, z=z, g=g, c=cbuf, sm=sm, sr=sr, s=s, x3=x3, u4=u4, m4=m4, r4=r4,
                         h2pre=h2pre, a2=a2, x4=x4, m5=m5, r5=r5, key_len=key_len, B=B, T=T, drop=drop, dsites=dsites))
    return y


def layer_backward(P, W: Operands, G: Dict[str, Tensor], pre: str, t: dict, dY: Tensor, cfg: Config,
                   sq: Optional[SideQueue] = None) -> Tensor:
    """dY: fp32 grad wrt the layer output -> fp32 grad wrt the layer input (dY's buffer is reused)."""
    B, T = t["B"], t["T"]
    N = B * T
    dev = dY.device
    TD = cfg.act_dtype
    f32 = torch.float32
    c = pre + "conv_module."

    bf16 = cfg.precision == "bf16"
    drop = t.get("drop")     # the forward's dropout site base: masks are regenerated from it (LayerNorm backward, fp32 path) ...
    dsites = t.get("dsites") or {}   # ... or read back as the forward's keep-mask words (tensor-core dSiLU and attention backward)

    def ln_bwd(dy, x_in, m, r, key, accumulate, want_h=True, bias_key=None, bias_scale=1.0, s_out=None):
        """LayerNorm backward into the residual-gradient stream dX; also emits the GEMM-operand copy of dX and,
        fused, the column sums of the new dX (= bias gradient `bias_key` of the next projection in the chain).
        s_out: dropout site of that projection's output -- copy and column sums then carry dropout'(dX)."""
        dsite = _at(drop, s_out) if s_out is not None else None
        dXh = _empty((N, D), TD, dev) if ((bf16 or dsite is not None) and want_h) else None
        ops.layernorm_bwd(dy, x_in, m, r, P[key + "weight"], dX, accumulate, G[key + "weight"], G[key + "bias"], dXh,
                          G[bias_key] if bias_key else None, bias_scale, drop=dsite)
        return dXh if dXh is not None else dX

    # final LayerNorm
    dX = _empty((N, D), f32, dev)
    dXh = ln_bwd(dY, t["x4"], t["m5"], t["r5"], pre + "final_layer_norm.", False, True, pre + "ffn2.sequential.4.bias", 0.5,
                 s_out=S_FFN2_OUT)

    def ffn_bwd(tag, x_in, u, m, r, hpre, a, dXh, want_h, s_act, next_bias=None, next_site=None):
        q = pre + tag + ".sequential."
        W1 = W.get(q + "1.weight", P[q + "1.weight"], (F, D))
        W2 = W.get(q + "4.weight", P[q + "4.weight"], (D, F))
        wgrad(dXh, a, G[q + "4.weight"], N, D, F, alpha=0.5, sq=sq)     # (this FFN's output-bias grad came fused from ln_bwd)
        dh = _empty((N, F), TD, dev)
        dgrad(dXh, W2, dh, N, D, F, act=ACT_DSILU, preact=hpre, alpha=0.5, drop=dsites.get(s_act, _at(drop, s_act)))
        wgrad(dh, u, G[q + "1.weight"], N, F, D, dbias=G[q + "1.bias"], sq=sq)
        du = _empty((N, D), TD if _BF16_DU else f32, dev)   # bf16 path: the gradient entering the LayerNorm backward is stored in bf16
        dgrad(dh, W1, du, N, F, D)
        return ln_bwd(du, x_in, m, r, q + "0.", True, want_h, next_bias, s_out=next_site)

    # FFN2: x4 = x3 + 0.5*FFN(u4), u4 = LN(x3)
    dXh = ffn_bwd("ffn2", t["x3"], t["u4"], t["m4"], t["r4"], t["h2pre"], t["a2"], dXh, True, S_FFN2_ACT, c + "sequential.5.bias",
                  S_CONV_OUT)

    kick = getattr(cfg, "_dp_kick", None)
    if kick is not None:
        kick()     # data parallel: gradient slices that are final leave for their all-reduce here (see OverlappedGradReducer.kick)

    # conv module: x3 = x2 + pw2(s) + b
    Wp1 = W.get(c + "sequential.0.weight", P[c + "sequential.0.weight"], (2 * D, D))
    Wp2 = W.get(c + "sequential.5.weight", P[c + "sequential.5.weight"], (D, D))
    wgrad(dXh, t["s"], G[c + "sequential.5.weight"].view(D, D), N, D, D, sq=sq)
    ds = _empty((N, D), TD, dev)
    dgrad(dXh, Wp2, ds, N, D, D)
    dc = _empty((N, D), f32, dev)
    zp = getattr(cfg, "_zero_pool", None)
    sums2 = zp.take() if zp is not None else torch.zeros(2 * D, dtype=torch.float64, device=dev)
    ops.bn_silu_bwd(ds, t["c"], t["sm"], t["sr"], P[c + "sequential.3.weight"], P[c + "sequential.3.bias"], sums2, dc,
                    G[c + "sequential.3.weight"], G[c + "sequential.3.bias"], sync=cfg.bn_sync, world=cfg.bn_world)
    dg = _empty((N, D), TD, dev)
    wdw = P[c + "sequential.2.weight"].detach().reshape(D, KW)
    ops.dwconv_bwd(dc, t["g"], wdw, dg, G[c + "sequential.2.weight"].view(D, KW), G[c + "sequential.2.bias"], B, T, KW)
    dz = _empty((N, 2 * D), TD, dev)
    ops.glu_bwd(t["z"], dg, dz)
    wgrad(dz, t["u3"], G[c + "sequential.0.weight"].view(2 * D, D), N, 2 * D, D, dbias=G[c + "sequential.0.bias"], sq=sq)
    du3 = _empty((N, D), TD if _BF16_DU else f32, dev)
    dgrad(dz, Wp1, du3, N, 2 * D, D)
    dXh = ln_bwd(du3, t["x2"], t["m3"], t["r3"], c + "layer_norm.", True, True, pre + "self_attn.out_proj.bias", s_out=S_ATTN_OUT)

    # MHSA: x2 = x1 + out_proj(attn(qkv)) ; qkv = in_proj(u2); u2 = LN(x1)
    Wqkv = W.get(pre + "self_attn.in_proj_weight", P[pre + "self_attn.in_proj_weight"], (3 * D, D))
    Wo = W.get(pre + "self_attn.out_proj.weight", P[pre + "self_attn.out_proj.weight"], (D, D))
    wgrad(dXh, t["ctx"], G[pre + "self_attn.out_proj.weight"], N, D, D, sq=sq)
    dctx = _empty((N, D), TD, dev)
    dgrad(dXh, Wo, dctx, N, D, D)
    dqkv = _empty((N, 3 * D), TD, dev)
    dvec = _empty((B * H * T,), f32, dev)
    dq32 = _empty((N, D), f32, dev) if bf16 else None
    ops.attn_bwd(t["qkv"], t["ctx"], dctx, t["lse"], t["key_len"], dqkv, dvec, B, T, H, dq32, drop=dsites.get(S_ATTN_P, _at(drop, S_ATTN_P)))
    wgrad(dqkv, t["u2"], G[pre + "self_attn.in_proj_weight"], N, 3 * D, D, dbias=G[pre + "self_attn.in_proj_bias"], sq=sq)
    du2 = _empty((N, D), TD if _BF16_DU else f32, dev)
    dgrad(dqkv, Wqkv, du2, N, 3 * D, D)
    dXh = ln_bwd(du2, t["x1"], t["m2"], t["r2"], pre + "self_attn_layer_norm.", True, True, pre + "ffn1.sequential.4.bias", 0.5,
                 s_out=S_FFN1_OUT)

    # FFN1
    ffn_bwd("ffn1", t["x"], t["u1"], t["m1"], t["r1"], t["h1pre"], t["a1"], dXh, False, S_FFN1_ACT)
    return dX


# ----------------------------------------------------------------------------------------------
# front end (early_exit.py:24-48 + positional_encoding.py:70-72)
# ----------------------------------------------------------------------------------------------
def frontend_forward(P, W: Operands, src: Tensor, cfg: Config, tape: Optional[dict], drop: Optional[ops.Drop] = None):
    B, n_mels, T_in = src.shape
    T1 = (T_in - 3) // 2 + 1
    T = (T1 - 3) // 2 + 1
    if T < 1:
        raise EecError(f"input too short: T_in={T_in}")
    pe = P["positional_encoder.pe"]
    if T > pe.shape[0]:
        raise RuntimeError(f"The size of tensor a ({T}) must match the size of tensor b ({pe.shape[0]}) at non-singleton dimension 0")
    dev, TD, f32 = src.device, cfg.act_dtype, torch.float32
    W1 = W.front_w1(P["conv_subsample.sequential.0.weight"])
    W2 = W.get("conv_subsample.sequential.1.weight", P["conv_subsample.sequential.1.weight"], (D, 3 * D))
    cols1 = _empty((B * T1, 256), TD, dev)
    ops.im2col_k3s2(src, n_mels * T_in, T_in, 1, cols1, 256, B, n_mels, T1)
    x1 = _empty((B * T1, D), f32, dev)
    linear(cols1, W1, x1, B * T1, D, 256, bias=P["conv_subsample.sequential.0.bias"])
    cols2 = _empty((B * T, 3 * D), TD, dev)
    ops.im2col_k3s2(x1, T1 * D, 1, D, cols2, 3 * D, B, D, T)
    x0 = _empty((B * T, D), f32, dev)
    linear(cols2, W2, x0, B * T, D, 3 * D, bias=P["conv_subsample.sequential.1.bias"], residual=pe.view(-1, D), res_row_mod=T)
    if drop is not None:
        ops.dropout(x0, x0, drop.at(SITE_PE))   # positional_encoding.py:72: dropout AFTER the sum with pe
    if tape is not None:
        tape.update(dict(cols1=cols1, cols2=cols2, B=B, T=T, T1=T1, n_mels=n_mels, drop=drop))
    return x0, T


def frontend_backward(P, W: Operands, G, t: dict, dX: Tensor, cfg: Config):
    B, T, T1, n_mels = t["B"], t["T"], t["T1"], t["n_mels"]
    dev, f32 = dX.device, torch.float32
    W2 = W.get("conv_subsample.sequential.1.weight", P["conv_subsample.sequential.1.weight"], (D, 3 * D))
    if t.get("drop") is not None:
        ops.dropout(dX, dX, t["drop"].at(SITE_PE))
    ops.colsum(dX, G["conv_subsample.sequential.1.bias"], B * T, D)
    dXh = to_act(dX, cfg)
    wgrad(dXh, t["cols2"], G["conv_subsample.sequential.1.weight"].view(D, 3 * D), B * T, D, 3 * D)
    dcols2 = _empty((B * T, 3 * D), f32, dev)
    dgrad(dXh, W2, dcols2, B * T, D, 3 * D)
    dx1 = _empty((B * T1, D), f32, dev)
    ops.col2im_k3s2(dcols2, 3 * D, dx1, B, D, T1, T)
    ops.colsum(dx1, G["conv_subsample.sequential.0.bias"], B * T1, D)
    dW1p = torch.zeros(256, 256, dtype=f32, device=dev)
    wgrad(to_act(dx1, cfg), t["cols1"], dW1p, B * T1, D, 256)
    k = 3 * n_mels
    G["conv_subsample.sequential.0.weight"].view(D, k).copy_(dW1p[:, :k])  # un-pad into the zeroed grad (layout plumbing)


# ----------------------------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------------------------
@dataclass
class Tape:
    front: dict = field(default_factory=dict)
    layers: List[dict] = field(default_factory=list)       # in execution order, each with "pre"
    heads: List[dict] = field(default_factory=list)
    branches: List[Optional[dict]] = field(default_factory=list)
    out: Optional[Tensor] = None
    B: int = 0
    T: int = 0


def check_lengths(lengths: Tensor, T: int):
    """torchaudio builds the key-padding mask with width max(length) (TA:11-14) and
    nn.MultiheadAttention asserts its shape: preserve that error."""
    if not lengths.is_cuda:
        mx = int(torch.clamp(lengths / 4, max=T).to(torch.int).max())
        if mx < T:
            raise AssertionError(f"Expected key_padded_mask.shape[1] to be {T}, but got {mx}")


def model_forward(P: Dict[str, Tensor], W: Operands, cfg: Config, src: Tensor, lengths: Tensor, training: bool,
                  want_tape: bool, side: Optional[dict] = None, drop_state: Optional[Tensor] = None):
    """-> (out [E,B,T,V] fp32 log-probs, Tape|None).  `side` (optional dict) receives per-exit
    argmax / frame-entropy tensors when it contains the key "want", and -- when it contains the key
    "want_hidden" -- side["hidden"] = [E,B,T,D] fp32 encoder states after every exit group (what
    full_conformer's decoders attend to, early_exit.py:783-786).
    drop_state: int64[2] device tensor {seed, offset} owned by THIS forward (train mode with cfg.drop_p > 0); the tape
    keeps it so that backward regenerates the same masks."""
    if not src.is_cuda:
        raise EecError("eec: input must be on a CUDA device; there is no CPU path")
    src = src.contiguous()
    if src.dtype != torch.float32:
        src = src.float()
    dev, f32 = src.device, torch.float32
    tape = Tape() if want_tape else None
    if training:
        cfg._zero_pool = _ZeroPool(cfg.n_exits * cfg.n_layers + 2, dev)
    drop0 = None
    if training and cfg.drop_p > 0.0:
        if drop_state is None:
            raise EecError("eec: train-mode forward with drop_prob > 0 needs a dropout state (seed, offset) tensor")
        drop0 = ops.Drop(drop_state, cfg.drop_p, 0)

    def layer_drop(uid: int):
        return drop0.at((1 + uid) * SITES_PER_LAYER) if drop0 is not None else None

    plan = None
    if drop0 is not None and cfg.precision == "bf16" and src.shape[2] >= 7:
        Tp = ((src.shape[2] - 3) // 2 + 1 - 3) // 2 + 1
        order = []
        for e in range(cfg.n_exits):
            order += [(e * cfg.n_layers + l, src.shape[0], Tp) for l in range(cfg.n_layers)]
            if cfg.splitformer and (e == 0 or e == cfg.last_exit):
                order.append((1000 + e // cfg.last_exit, src.shape[0], (Tp + 1) // 2))
        plan = MaskPlan(drop0, order)

    x, T = frontend_forward(P, W, src, cfg, tape.front if tape else None, drop0)
    B = src.shape[0]
    N = B * T
    check_lengths(lengths, T)
    lengths_dev = lengths.to(device=dev, dtype=torch.int64, non_blocking=True)
    key_len = _empty((B,), torch.int32, dev)
    ops.encoder_lengths(lengths_dev, key_len, T, 4, 0)
    E = cfg.n_exits
    out = _empty((E, B, T, V), f32, dev)
    logits_ws = _empty((N, V), f32, dev) if cfg.precision == "fp32" else None
    want_side = side is not None and "want" in side
    hidden = None
    if side is not None and "want_hidden" in side:
        hidden = _empty((E, B, T, D), f32, dev)
        side["hidden"] = hidden
    for e in range(E):
        x_in = x
        for l in range(cfg.n_layers):
            pre = f"conformer.{e}.conformer_layers.{l}."
            lt = {"pre": pre} if tape else None
            uid = e * cfg.n_layers + l
            x = layer_forward(P, W, pre, x, key_len, B, T, cfg, training, lt, layer_drop(uid), plan.take(uid) if plan else None)
            if tape:
                tape.layers.append(lt)
        br = None
        if cfg.splitformer and (e == 0 or e == cfg.last_exit):
            # early_exit.py:314-356: parallel stride-2 branch on the group's INPUT; its key mask uses the RAW
            # fbank lengths (reference quirk) -> clamp((lengths+pad)/2, max=T2)
            i = e // cfg.last_exit     # (a 1-exit Splitformer divides by zero here exactly like the reference, early_exit.py:316)
            T2 = (T + 1) // 2
            pad = T % 2
            xd = _empty((B * T2, D), f32, dev)
            ops.stride2_gather(x_in, xd, B, T)
            len2 = _empty((B,), torch.int32, dev)
            ops.encoder_lengths(lengths_dev, len2, T2, 2, pad)
            pre = f"conformer_parallel.{i}.conformer_layers.0."
            br = {"pre": pre} if tape else None
            yd = layer_forward(P, W, pre, xd, len2, B, T2, cfg, training, br, layer_drop(1000 + i), plan.take(1000 + i) if plan else None)
            if x is x_in:
                x = x.clone()
            ops.repeat2_add(yd, x, B, T)
        if tape:
            tape.branches.append(br)
        if hidden is not None:
            hidden[e].view(N, D).copy_(x)   # layout plumbing: the residual stream buffer is reused by later layers
        xh = to_act(x, cfg)
        Wh = W.get(f"linears.{e}.weight", P[f"linears.{e}.weight"], (V, D))
        am = _empty((N,), torch.int32, dev) if want_side else None
        en = _empty((N,), f32, dev) if want_side else None
        ops.head_logsoftmax(xh, Wh, P[f"linears.{e}.bias"], out[e], am, en, logits_ws)
        if want_side:
            side.setdefault("argmax", []).append(am.view(B, T))
            side.setdefault("entropy", []).append(en.view(B, T))
        if tape:
            tape.heads.append({"xh": xh})
    if plan is not None:
        plan.join()
    if want_side:
        side["key_len"] = key_len
    if tape:
        tape.out, tape.B, tape.T = out, B, T
    return out, tape


# Gradients that are known to be d(loss)/d(logits) already.  The fused CTC kernel returns (softmax - occupancy) scaled per utterance:
# its rows sum to zero, so the log-softmax backward (g - softmax * rowsum(g)) would hand it back unchanged.  _CtcFn.backward registers the
# tensor it returns; when exactly that tensor (same storage, shape and version) arrives as the encoder's upstream gradient, the six
# log-softmax backward launches are skipped.  Anything else (a user's own loss on the log-probs, an accumulated gradient) misses the
# registry and takes the general path.  An entry is only valid inside the backward pass (autograd graph task) that created it.
_LOGIT_GRADS: Dict[int, tuple] = {}


def _backward_pass_id() -> int:
    fn = getattr(torch._C, "_current_graph_task_id", None)
    return int(fn()) if fn is not None else -1


def mark_logit_grad(g: Tensor) -> None:
    _LOGIT_GRADS.clear()                      # (one pending step at a time: nothing accumulates here)
    task = _backward_pass_id()
    if task >= 0:
        _LOGIT_GRADS[g.data_ptr()] = (tuple(g.shape), g._version, task)


def take_logit_grad(g: Tensor) -> bool:
    hit = _LOGIT_GRADS.pop(g.data_ptr(), None)
    return hit is not None and hit == (tuple(g.shape), g._version, _backward_pass_id())


def group_ranges(P, names: List[str], n_exits: int):
    """[lo, hi) of every exit group's parameters inside the flat gradient buffer (layout = `names` order), or None when a
    group is not contiguous there.  Backward finishes the groups last-to-first: each slice can leave for the data-parallel
    all-reduce while the earlier groups are still being differentiated."""
    off, spans = 0, {}
    for n in names:
        k = P[n].numel()
        if n.startswith("conformer."):
            e = int(n.split(".")[1])
            lo, hi = spans.get(e, (off, off))
            if hi != off:
                return None
            spans[e] = (lo, off + k)
        off += k
    if sorted(spans) != list(range(n_exits)):
        return None
    return [spans[e] for e in range(n_exits)], off


def model_backward(P, W: Operands, cfg: Config, tape: Tape, gout: Tensor, names: List[str],
                   ghid: Optional[Tensor] = None, on_ready=None) -> Dict[str, Tensor]:
    """gout: grad wrt out [E,B,T,V] (fp32); ghid (optional): grad wrt the per-exit encoder states [E,B,T,D]
    (the decoders' cross-attention in AED mode).  Returns fp32 grads for every name in `names`.
    on_ready(flat, lo, hi) (optional) is called as soon as flat[lo:hi] holds final gradients (one exit group at a time,
    then the rest): the data-parallel reducer's hook (eec.distributed.OverlappedGradReducer)."""
    dev, f32 = gout.device, torch.float32
    logit_grad = gout.is_contiguous() and take_logit_grad(gout)
    gout = gout.contiguous()
    B, T = tape.B, tape.T
    N = B * T
    E = cfg.n_exits
    # all parameter gradients live in ONE flat fp32 buffer (one memset; one NCCL all-reduce under DP)
    total = sum(P[n].numel() for n in names)
    cfg._zero_pool = _ZeroPool(cfg.n_exits * cfg.n_layers + 2, dev)
    flat = torch.zeros(total, dtype=f32, device=dev)
    G, off = {}, 0
    for n in names:
        k = P[n].numel()
        G[n] = flat[off:off + k].view(P[n].shape)
        off += k
    G["__flat__"] = flat
    spans = group_ranges(P, names, E) if on_ready is not None else None
    sq = SideQueue(dev) if (_WGRAD_SIDE and cfg.precision == "bf16") else None   # weight gradients run beside the data-gradient chain
    dX: Optional[Tensor] = None
    li = len(tape.layers)
    for e in reversed(range(E)):
        # exit head: out[e] = log_softmax(x W^T + b)
        if logit_grad:
            dlog = gout[e].view(N, V)         # already d/d(logits) (fused CTC): log-softmax backward is the identity on it
        else:
            dlog = _empty((N, V), f32, dev)
            ops.logsoftmax_bwd(gout[e], tape.out[e], dlog)
        Wh = W.get(f"linears.{e}.weight", P[f"linears.{e}.weight"], (V, D))
        if cfg.precision == "bf16":
            dlogh = _empty((N, V), torch.bfloat16, dev)
            ops.cast_colsum(dlog, dlogh, G[f"linears.{e}.bias"], N, V)     # operand copy + bias gradient in one pass
        else:
            dlogh = dlog
            ops.colsum(dlog, G[f"linears.{e}.bias"], N, V)
        wgrad(dlogh, tape.heads[e]["xh"], G[f"linears.{e}.weight"], N, V, D, sq=sq)
        if dX is None:
            dX = _empty((N, D), f32, dev)
            dgrad(dlogh, Wh, dX, N, V, D)
        else:
            dgrad(dlogh, Wh, dX, N, V, D, residual=dX)
        if ghid is not None:
            dX.add_(ghid[e].reshape(N, D))   # plumbing: the second consumer of this exit's state
        br = tape.branches[e]
        d_in_extra = None
        if br is not None:
            # x = x_main + repeat2(yd): d(yd)[t2] = dX[2 t2] + dX[2 t2 + 1]; d(x_in) += scatter(d(xd))
            T2 = (T + 1) // 2
            dyd = _empty((B * T2, D), f32, dev)
            ops.repeat2_bwd(dX, dyd, B, T)
            d_in_extra = layer_backward(P, W, G, br["pre"], br, dyd, cfg, sq)
        for l in reversed(range(cfg.n_layers)):
            li -= 1
            lt = tape.layers[li]
            dX = layer_backward(P, W, G, lt["pre"], lt, dX, cfg, sq)
        if d_in_extra is not None:
            ops.stride2_scatter_add(d_in_extra, dX, B, T)
        if spans is not None:
            if sq is not None:
                sq.join()          # this group's weight gradients are final before its slice leaves for the all-reduce
            on_ready(flat, *spans[0][e])
    frontend_backward(P, W, G, tape.front, dX, cfg)
    if sq is not None:
        sq.join()
    if on_ready is not None:
        if spans is None:
            on_ready(flat, 0, total)
        else:   # what is not an exit group: front end + heads (before the groups), Splitformer branches (after)
            lo0, hi0 = spans[0][0][0], spans[0][-1][1]
            if lo0 > 0:
                on_ready(flat, 0, lo0)
            if hi0 < total:
                on_ready(flat, hi0, total)
    return G
