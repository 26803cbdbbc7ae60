"""Optimiser step of the reference's training loop over flat buffers.

``FusedNoamAdamW`` replaces, with three kernel launches and no host sync,

    torch.nn.utils.clip_grad_norm_(model.parameters(), args.clip)                      # train.py:69
    NoamOpt(d_model, warmup, AdamW(model.parameters(), lr=0, betas=(0.9, 0.98),        # train.py:261-262
            eps=args.adam_eps, weight_decay=args.weight_decay)).step()                  # util/noam_opt.py:26-40

The engine's backward already leaves every gradient in one flat fp32 buffer (``model._flat_grad``); the constructor lays
the parameters out the same way (each ``p.data`` becomes a view of one flat buffer: state_dict, checkpoints and
``model.parameters()`` are unchanged) together with the two Adam moments, so the update is one bandwidth-bound pass that
also emits the bf16 GEMM-operand copy of every weight for the next forward (no per-tensor cast launches).  Step counter,
gradient norm and learning rate live on the device, so the step can be captured in the training CUDA graph
(``GraphedTrainStep(..., optimizer=opt)``).  Unlike NoamOpt.step() it neither prints the rate nor syncs (noam_opt.py:33).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .lib import EecError, on_device


class FusedNoamAdamW:
    def __init__(self, model, model_size: int = 256, warmup: int = 25000, betas=(0.9, 0.98), eps: float = 1e-9,
                 weight_decay: float = 5e-4, clip: float = 1.0, lr: float | None = None):
        params = list(model.named_parameters())
        if not params or not params[0][1].is_cuda:
            raise EecError("FusedNoamAdamW: move the model to a CUDA device first (no CPU path)")
        if any(p.dtype != torch.float32 for _, p in params):
            raise EecError("FusedNoamAdamW: parameters must be fp32 (bf16 operand copies are made by the step itself)")
        if hasattr(model, "_encoder_named_parameters"):
            raise EecError("FusedNoamAdamW covers models whose every gradient comes from the eec engine "
                           "(Early_conformer / Splitformer); use torch.optim for full_conformer's decoder half")
        self.model = model
        self.model_size, self.warmup, self.betas, self.eps = float(model_size), float(warmup), betas, float(eps)
        self.weight_decay, self.clip, self.lr = float(weight_decay), float(clip), lr
        dev = params[0][1].device
        self.total = sum(p.numel() for _, p in params)
        self.flat_p = torch.empty(self.total, dtype=torch.float32, device=dev)
        self.offsets: Dict[str, tuple] = {}
        off = 0
        with torch.no_grad():
            for n, p in params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)     # parameters become views of the flat buffer
                self.offsets[n] = (off, k)
                off += k
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.state = torch.zeros(4, dtype=torch.float64, device=dev)   # step, sum g^2, lr, clip coefficient
        self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=dev)
        with on_device(self.flat_p.device):
            ops.cast(self.flat_p, self.shadow)
        self._versions = {n: p._version for n, p in params}
        self._attach()
        model.__dict__["_fused_optimizer"] = self     # (eec.distributed.broadcast_parameters refreshes the shadow through this)

    def _attach(self):
        ob = self.model._operands
        ob.attach_shadow(self.shadow, self.offsets, self._versions)

    def step(self) -> None:
        """Call after ``loss.backward()`` (and after the DP all-reduce).  Asynchronous on the current stream."""
        m = self.model
        flat_g = m.__dict__.get("_flat_grad")
        first = next(m.parameters())
        if flat_g is None or first.grad is None or first.grad.data_ptr() != flat_g.data_ptr() or flat_g.numel() != self.total:
            raise EecError("FusedNoamAdamW.step: gradients are not in the engine's flat buffer (run loss.backward() on an eec model first)")
        if first.data_ptr() != self.flat_p.data_ptr():
            raise EecError("FusedNoamAdamW.step: parameters were re-allocated (model.to()/load with assign=True?) after the optimiser was built")
        with on_device(self.flat_p.device):
            ops.call("eec_noam_adamw_step", ops.ptr(self.flat_p), ops.ptr(flat_g), ops.ptr(self.exp_avg), ops.ptr(self.exp_avg_sq),
                     ops.ptr(self.shadow), self.total, ops.ptr(self.state), self.model_size, self.warmup, float(self.betas[0]),
                     float(self.betas[1]), self.eps, self.weight_decay, self.clip, -1.0 if self.lr is None else float(self.lr), ops.stream())
        ob = m._operands
        if ob._shadow is not self.shadow:
            self._attach()          # precision switch rebuilt the operand cache
        ob.invalidate()             # cached casts of re-laid-out operands (front-end weight) are stale now

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.model.zero_grad(set_to_none=set_to_none)

    # ---- introspection / checkpointing (these read device state: they sync) ----------------------------------
    @property
    def _step(self) -> int:
        return int(self.state[0].item())

    def rate(self, step: int | None = None) -> float:
        """util/noam_opt.py:35-40."""
        if self.lr is not None:
            return float(self.lr)
        step = self._step if step is None else step
        return self.model_size ** (-0.5) * min(step ** (-0.5), step * self.warmup ** (-1.5)) if step > 0 else 0.0

    def last_grad_norm(self) -> float:
        return float(self.state[1].item()) ** 0.5

    def state_dict(self):
        """A superset of the reference's NoamOpt.state_dict() (util/noam_opt.py:12-17: every attribute but the wrapped optimiser, i.e.
        `_step`, `warmup`, `model_size`, `_rate`), so a file written here loads into the reference's NoamOpt (`__dict__.update`, :19-25:
        the extra keys are inert attributes there), plus the Adam moments the reference never saves (train.py:122-125 stores the model
        and this scheduler state only, so a resumed reference run restarts its moments at zero)."""
        step = self._step
        return {"_step": step, "warmup": self.warmup, "model_size": self.model_size, "_rate": self.rate(step),
                "step": step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        """Accepts this class's files and the reference's `lr###-transformer` NoamOpt state (`_step`, `warmup`, `model_size`, `_rate`);
        moments missing from the file start at zero, which is exactly what the reference does on resume."""
        step = sd["_step"] if "_step" in sd else sd["step"]
        self.state.zero_()
        self.state[0] = float(step)
        if "warmup" in sd:
            self.warmup = float(sd["warmup"])
        if "model_size" in sd:
            self.model_size = float(sd["model_size"])
        for name, buf in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
            if sd.get(name) is not None:
                buf.copy_(sd[name])
            else:
                buf.zero_()

    def refresh_shadow(self) -> None:
        """Call after writing parameters behind the optimiser's back (e.g. load_state_dict into the model)."""
        with on_device(self.flat_p.device):
            ops.cast(self.flat_p, self.shadow)
        self._versions = {n: p._version for n, p in self.model.named_parameters()}
        self._attach()
        self.model._operands.invalidate()
