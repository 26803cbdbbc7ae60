"""Batch-sharded data parallelism for the CTC training step.

The reference has no distributed code at all; BASELINE configs[2] asks for one rank per
GPU of an 8xB200 box with the gradient all-reduce over NCCL/NVLink.  The path shards by utterance:
every rank runs the full model on its own 64-utterance batch (weak scaling), gradients are averaged.

`eec.engine.model_backward` writes every parameter gradient into ONE flat fp32 buffer (the `p.grad`
tensors are views of it), so the exchange step is a single `all_reduce` of 126 MB (12L) / 188 MB (18L)
instead of 413 / 611 small ones.  BatchNorm batch statistics stay per-rank (like torch DDP's default);
the reference defines no multi-GPU semantics to match.

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


def all_reduce_gradients(model: torch.nn.Module, average: bool = True) -> None:
    """Call after `loss.backward()`.  Uses the engine's flat buffer when present (one collective); falls back
    to flattening `p.grad` (e.g. for a model that was not run through eec in this step)."""
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
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)
