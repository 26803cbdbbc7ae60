"""Batch-sharded data parallelism for the CTC training step.

The reference has no distributed code at all; BASELINE configs[2] asks for one rank per
GPU of an 8xB200 box with the gradient all-reduce over NCCL/NVLink.  The path shards by utterance:
every rank runs the full model on its own 64-utterance batch (weak scaling), gradients are averaged.

`eec.engine.model_backward` writes every parameter gradient into ONE flat fp32 buffer (the `p.grad`
tensors are views of it), so the exchange step is a single `all_reduce` of 126 MB (12L) / 188 MB (18L)
instead of 413 / 611 small ones.  BatchNorm batch statistics stay per-rank by default (like torch DDP's default; the
reference defines no multi-GPU semantics to match); `sync_batchnorm(model)` switches the 12 / 18 BatchNorm layers to GLOBAL
batch statistics (two 4 KB SUM all-reduces per layer and step), which makes an N-rank step on N shards of a batch compute
exactly the 1-rank step on the concatenated batch -- same loss, same averaged gradients, same running statistics
(`tools/dp_check.py` checks that on real GPUs).  Without it, `sync_bn_buffers(model)` averages the per-rank running
statistics before a checkpoint is written.

Works with any torch.distributed backend: `nccl` on the GPU box, `gloo` in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_items` utterances owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    """Inference / training sharding: slice every per-utterance tensor along dim 0; no collective."""
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def all_reduce_flat(flat: torch.Tensor, average: bool = True) -> torch.Tensor:
    """In-place all-reduce of a flat gradient buffer (AVG where the backend has it, else SUM / world)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    if average and dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(dist.get_world_size())
    return flat


class OverlappedGradReducer:
    """All-reduce of the gradients DURING backward, one exit group at a time, on a side stream.

        red = eec.distributed.OverlappedGradReducer(model)      # once, after dist.init_process_group
        loss.backward()                                         # gradients are already averaged when this returns
        optimizer.step()

    The engine's backward finishes the exit groups last-to-first and calls `on_ready(flat, lo, hi)` when a group's slice of the
    flat fp32 gradient buffer is final (12 layers / 6 exits: 21 MB per group, then 1.3 MB of front end + heads): the slice is
    all-reduced over NCCL on `self.stream` while the earlier groups are still being differentiated on the compute stream;
    `finish()` (called by the engine at the end of backward) joins the two streams, so `clip_grad_norm_` / the optimiser see
    the complete, averaged buffer.  Everything is stream-ordered with no host synchronisation, so `GraphedTrainStep` captures
    the collectives as parallel branches of the step's CUDA graph: a data-parallel step is still ONE graph launch.
    BatchNorm statistics stay per rank.  With CPU tensors (gloo tests) the reduction is issued in line."""

    def __init__(self, model: torch.nn.Module, average: bool = True, group=None, grad_dtype: str = "fp32"):
        """grad_dtype "bf16": every slice is cast to a bf16 staging buffer, all-reduced in bf16 (half the NVLink bytes: 63 MB instead of
        126 MB at 12 layers) and cast back -- an 8-bit-mantissa exchange, opt-in; the flat buffer and the optimiser stay fp32."""
        if grad_dtype not in ("fp32", "bf16"):
            raise ValueError(f"grad_dtype must be 'fp32' or 'bf16', got {grad_dtype!r}")
        self.average, self.group, self.grad_dtype = average, group, grad_dtype
        self._stage = {}
        p = next(model.parameters())
        self.stream = torch.cuda.Stream(device=p.device) if p.is_cuda else None
        self._forked = False
        self.calls = 0            # slices reduced since construction (tests / bench report it)
        # EEC_DP_DEFER=1 (opt-in): a slice that is final is not all-reduced at once but at the next "kick" of backward (before the
        # convolution module's backward of the following layer): the NCCL kernel holds SMs for ~140 us at N = 8, and a persistent
        # one-CTA-per-SM GEMM (148 CTAs, static work split) that cannot be fully resident finishes a wave late, whereas the convolution /
        # attention backward kernels that follow the kick are grid-strided or dynamically scheduled and simply share the SMs.
        # Measured on 2 x B200: 12.29 ms per step against 12.28-12.38 ms without (neutral; gradients identical, tools/dp_check.py);
        # not measured at N = 8, hence off by default.
        import os
        self.defer = os.environ.get("EEC_DP_DEFER", "0") == "1"
        self._pending = []
        model.__dict__["_grad_reducer"] = self

    def _reduce(self, t: torch.Tensor) -> None:
        if self.grad_dtype == "bf16" and t.is_cuda:
            from . import ops
            st = self._stage.get((t.data_ptr(), t.numel()))
            if st is None:
                st = self._stage[(t.data_ptr(), t.numel())] = torch.empty(t.numel(), dtype=torch.bfloat16, device=t.device)
            ops.cast(t, st)
            dist.all_reduce(st, op=dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM, group=self.group)
            ops.cast(st, t)
            return
        if self.average and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                t.div_(dist.get_world_size(self.group))

    def on_ready(self, flat: torch.Tensor, lo: int, hi: int) -> None:
        if hi <= lo or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        self.calls += 1
        sl = flat[lo:hi]
        if self.stream is None or not flat.is_cuda:
            self._reduce(sl)
            return
        if self.defer:
            self._pending.append(sl)
            return
        self.stream.wait_stream(torch.cuda.current_stream(flat.device))   # the slice is final on the compute stream
        with torch.cuda.stream(self.stream):
            self._reduce(sl)
        self._forked = True

    def kick(self) -> None:
        """all-reduce the slices that became final since the last kick (called by the engine's backward at points where the kernels that
        follow tolerate sharing the SMs with the collective, and by finish())"""
        if not self._pending:
            return
        dev = self._pending[0].device
        self.stream.wait_stream(torch.cuda.current_stream(dev))   # everything issued so far, the slices' last writers included
        with torch.cuda.stream(self.stream):
            for sl in self._pending:
                self._reduce(sl)
        self._pending = []
        self._forked = True

    def finish(self) -> None:
        self.kick()
        if self._forked:
            torch.cuda.current_stream(self.stream.device).wait_stream(self.stream)
            self._forked = False

    def remove(self, model: torch.nn.Module) -> None:
        if model.__dict__.get("_grad_reducer") is self:
            del model.__dict__["_grad_reducer"]


def all_reduce_gradients(model: torch.nn.Module, average: bool = True) -> None:
    """Call after `loss.backward()`.  Uses the engine's flat buffer when present (one collective); falls back
    to flattening `p.grad` (e.g. for a model that was not run through eec in this step).  A no-op when an
    OverlappedGradReducer is installed on the model: backward has already averaged the gradients."""
    if model.__dict__.get("_grad_reducer") is not None:
        return
    flat = model.__dict__.get("_flat_grad")
    params = [p for p in model.parameters() if p.grad is not None]
    if flat is not None and params and params[0].grad.data_ptr() == flat.data_ptr():
        all_reduce_flat(flat, average)
        return
    if not params:
        return
    buf = torch.cat([p.grad.reshape(-1) for p in params])
    all_reduce_flat(buf, average)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(buf[off:off + n].view_as(p.grad))
        off += n


def broadcast_parameters(model: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers.  A FusedNoamAdamW built earlier on the model keeps a bf16
    operand shadow of the parameters: it is refreshed here (`.data` writes do not bump tensor versions, so nothing else would notice)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)
    opt = model.__dict__.get("_fused_optimizer")
    if opt is not None:
        opt.refresh_shadow()
    ob = model.__dict__.get("_operands_obj")
    if ob is not None:
        ob.invalidate()


class _BnSync:
    """SUM all-reduce of one BatchNorm layer's per-channel statistics (double[2*C]) across the data-parallel ranks, in place, on the
    compute stream (the statistics are on the critical path: the normalisation that follows needs them)."""

    def __init__(self, group=None):
        self.group = group
        self.calls = 0

    def __call__(self, sums: torch.Tensor) -> None:
        self.calls += 1
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)


def sync_batchnorm(model: torch.nn.Module, enable: bool = True, group=None) -> None:
    """Synchronised BatchNorm for data-parallel training: every conformer convolution module normalises with the statistics
    of the GLOBAL batch -- forward: the per-channel sum / sum of squares are all-reduced before the normalisation; backward: the two
    per-channel sums of the BatchNorm adjoint likewise (24 / 36 all-reduces of 4 KB per step at 12 / 18 layers, captured in the step's
    graph).  Every rank must run the same batch shape (B, T_in).  With it the N-rank step equals the 1-rank step on the concatenated
    batch, and the running statistics are identical on every rank."""
    if enable and dist.is_initialized() and dist.get_world_size(group) > 1:
        model.__dict__["_bn_sync"] = (_BnSync(group), dist.get_world_size(group))
    else:
        model.__dict__.pop("_bn_sync", None)


def sync_bn_buffers(model: torch.nn.Module, group=None) -> None:
    """Average the BatchNorm running statistics over the ranks (call before saving a checkpoint when BatchNorm is NOT synchronised:
    each rank has tracked the statistics of its own shards).  num_batches_tracked is the same on every rank already."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    w = dist.get_world_size(group)
    for n, b in model.named_buffers():
        if n.endswith("running_mean") or n.endswith("running_var"):
            dist.all_reduce(b.data, op=dist.ReduceOp.SUM, group=group)
            b.data.div_(w)
