"""CTC loss over libeec.so's fused forward-backward kernel.

``CTCLoss`` keeps the call shape of the ``torch.nn.CTCLoss(blank=0, zero_infinity=True)`` instance the
reference builds at train.py:259 and calls at train.py:61:
``ctc_loss(log_probs (T,B,V), targets (B,L), input_lengths (B), target_lengths (B)) -> scalar``.
``multi_exit_ctc_loss`` is the fused form of train.py:57-63 (all exits in ONE launch).
Like ATen, the gradient handed back for ``log_probs`` is ``(exp(lp) - occupancy) / (B * max(U_b, 1))``.
"""
from __future__ import annotations

import torch

from . import ops
from .lib import EecError, on_device


class _CtcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lp_ebtv, targets, target_lengths, blank):
        E, B, T, V = lp_ebtv.shape
        dev = lp_ebtv.device
        tg = targets.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        tl = target_lengths.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        nll = torch.empty(E, B, dtype=torch.float32, device=dev)
        loss = torch.zeros(E, dtype=torch.float32, device=dev)
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(lp_ebtv) if need_grad else None
        with on_device(dev):
            ops.ctc_fwd_bwd(lp_ebtv, tg, tl, nll, loss, grad, 1.0, blank)
        ctx.grad = grad
        ctx.nll = nll
        return loss

    @staticmethod
    def backward(ctx, gloss):
        g = ctx.grad
        if g is None:
            raise EecError("ctc: gradient was not computed in forward")
        gl = gloss.contiguous().float()
        with on_device(g.device):
            ops.scale_rows_dev(g, gl, g)   # every exit's slab times its upstream scalar, on device, one launch
        ctx.grad = None
        from . import engine
        engine.mark_logit_grad(g)      # rows sum to zero: the encoder's backward may skip the log-softmax backward for this tensor
        return g, None, None, None


def multi_exit_ctc_loss(out_ebtv: torch.Tensor, targets: torch.Tensor, target_lengths: torch.Tensor, blank: int = 0,
                        reduce_exits: bool = True):
    """sum_e mean_b( nll_{e,b} / max(U_b,1) ), zero_infinity=True, input length = T' for every utterance
    (train.py:57-63).  out_ebtv: (E,B,T',V) fp32 log-probs as returned by Early_conformer.forward."""
    if not out_ebtv.is_cuda:
        raise EecError("ctc: log-probs must be on a CUDA device (no CPU path)")
    if out_ebtv.dtype != torch.float32 or not out_ebtv.is_contiguous():
        raise EecError("ctc: expected contiguous fp32 (E,B,T,V) log-probs")
    per_exit = _CtcFn.apply(out_ebtv, targets, target_lengths, blank)
    return per_exit.sum() if reduce_exits else per_exit


class CTCLoss(torch.nn.Module):
    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = True):
        super().__init__()
        if reduction != "mean" or not zero_infinity:
            raise NotImplementedError("eec.CTCLoss implements the reference configuration only: reduction='mean', zero_infinity=True")
        self.blank = blank

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        if log_probs.dim() != 3:
            raise EecError("ctc: log_probs must be (T,B,V)")
        T, B, V = log_probs.shape
        if not input_lengths.is_cuda and not bool((input_lengths == T).all()):
            raise NotImplementedError("eec.CTCLoss: the reference always passes input_lengths == T' (train.py:57-58)")
        btv = log_probs.permute(1, 0, 2)
        if not btv.is_contiguous():
            btv = btv.contiguous()
        return _CtcFn.apply(btv.unsqueeze(0), targets, target_lengths, self.blank)[0]
