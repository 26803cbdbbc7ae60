"""Whole training step as ONE CUDA graph (north star: "CUDA streams and graphs instead of a tracing compiler").

The eager step (train.py:53-69: forward, summed multi-exit CTC, backward) issues ~600 kernel launches through
ctypes; at ~20 ms of GPU work per step the Python launch path is as long as the GPU time, so every kernel speed-up
disappears behind launch gaps.  ``GraphedTrainStep`` runs the step once under ``torch.cuda.graph`` with static
input / target buffers and replays it: one ``cudaGraphLaunch`` per step.

    step = eec.GraphedTrainStep(model, batch_size=64, t_in=1501, max_target_len=82)
    loss = step(src, lengths, targets, target_lengths)      # same tensors the reference's train() feeds
    eec.distributed.all_reduce_gradients(model)              # DP: one flat-buffer all-reduce (outside the graph)
    optimizer.step()

What is inside the graph: bf16 operand casts of every parameter (so weight updates between replays are seen),
the encoder forward, the fused 6-exit CTC forward-backward, the backward pass into ONE flat fp32 gradient buffer
(zeroed inside the graph).  ``p.grad`` tensors are views of that buffer and stay valid across replays.
What stays outside: host-side argument checks (the reference's ``max(lengths)//4 >= T'`` precondition,
early_exit.py:623 / TA:9-15), host->device copies of the step's inputs, the NCCL all-reduce, the optimiser.
Shapes are static: a new (batch, T_in, max_target_len) needs a new GraphedTrainStep (the reference's length-sorted
sub-batches, data_loader.py:166-188, map to a small set of bucketed shapes).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine
from .ctc import multi_exit_ctc_loss
from .lib import EecError, on_device, load


class GraphedTrainStep:
    def __init__(self, model, batch_size: int, t_in: int, max_target_len: int, n_mels: Optional[int] = None, blank: int = 0,
                 pad_token: int = 126, warmup: int = 2, optimizer=None):
        params = list(model.parameters())
        if not params or not params[0].is_cuda:
            raise EecError("GraphedTrainStep: move the model to a CUDA device first (no CPU path)")
        if not model.training:
            raise EecError("GraphedTrainStep captures a TRAINING step: call model.train() first")
        dev = params[0].device
        self.model, self.blank, self.pad_token = model, blank, pad_token
        # optimizer (eec.FusedNoamAdamW, single-GPU): its three launches join the graph, so a replay is a COMPLETE training
        # step; under data parallelism leave it None and call all_reduce_gradients(model); opt.step() after the replay
        self.optimizer = optimizer
        self.B, self.T_in, self.L = batch_size, t_in, max_target_len
        n_mels = n_mels if n_mels is not None else model._features_length
        # static device buffers the graph reads, and pinned host staging for the per-step uploads
        self.src = torch.zeros(batch_size, n_mels, t_in, dtype=torch.float32, device=dev)
        self.lengths = torch.full((batch_size,), t_in, dtype=torch.int64, device=dev)
        self.targets = torch.full((batch_size, max_target_len), pad_token, dtype=torch.int64, device=dev)
        self.target_lengths = torch.ones(batch_size, dtype=torch.int64, device=dev)
        self._pin_src = torch.empty(self.src.shape, dtype=torch.float32).pin_memory()
        self._pin_small = torch.empty(batch_size * (max_target_len + 2), dtype=torch.int64).pin_memory()
        self._dev_small = torch.empty_like(self._pin_small, device=dev)
        self.t_out = ((t_in - 3) // 2 + 1 - 3) // 2 + 1

        lib = load()
        self._uploaded = None            # event after the last H2D copy out of the pinned staging buffers
        # warm-up and capture run the step on the all-zero static input: they must leave NO trace in the model -- the BatchNorm
        # running statistics / num_batches_tracked and the dropout counter are put back afterwards (the weights do not move:
        # the optimiser is left out of the warm-up, and the captured graph is not executed by the capture)
        saved = {n: b.detach().clone() for n, b in model.named_buffers()}
        drop_state = model.__dict__.get("_drop_state")
        saved_drop = drop_state.clone() if drop_state is not None else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):   # eager warm-up off the default stream (lazy module loads, kernel attributes, tensor maps)
            for _ in range(max(warmup, 1)):
                self._eager_step(with_optimizer=False)   # (warm-up must not move the weights)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for n, b in model.named_buffers():
                b.copy_(saved[n])
            if model.__dict__.get("_drop_state") is not None:
                if saved_drop is not None:
                    model.__dict__["_drop_state"].copy_(saved_drop)
                else:
                    model.__dict__["_drop_state"][1] = 0
        # parameter casts must be IN the graph: drop the operand cache so capture re-issues them
        model._operands._cache.clear()
        model.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.eec_launch_count()
        with torch.cuda.graph(self.graph):
            loss = self._eager_step()
        self.launches_per_step = int(lib.eec_launch_count() - n0)   # kernels of ours inside one replay
        self.loss = loss.detach()
        self.flat_grad = model.__dict__.get("_flat_grad")

    def _eager_step(self, with_optimizer: bool = True):
        out = self.model(self.src, self.lengths)
        loss = multi_exit_ctc_loss(out, self.targets, self.target_lengths, self.blank)
        loss.backward()
        if with_optimizer and self.optimizer is not None:
            self.optimizer.step()
        return loss

    def load_inputs(self, src: torch.Tensor, lengths: torch.Tensor, targets: torch.Tensor, target_lengths: torch.Tensor) -> None:
        """Validate one batch on the host and upload it into the graph's static buffers (async on the current stream)."""
        B, L = self.B, self.L
        if tuple(src.shape) != tuple(self.src.shape):
            raise EecError(f"GraphedTrainStep: src shape {tuple(src.shape)} != captured {tuple(self.src.shape)}")
        if targets.shape[0] != B or targets.shape[1] > L or lengths.numel() != B or target_lengths.numel() != B:
            raise EecError("GraphedTrainStep: batch / target shape does not fit the captured step")
        if not lengths.is_cuda:
            engine.check_lengths(lengths, self.t_out)   # the reference's AssertionError, raised on the host before any launch
        # the pinned staging buffers are single: a previous call's asynchronous H2D copy may still be queued behind an earlier
        # replay, so wait for it before the host overwrites them (otherwise that step would train on a torn / future batch)
        self._wait_uploaded()
        if src.is_cuda:
            self.src.copy_(src, non_blocking=True)
        else:
            self._pin_src.copy_(src)
            self.src.copy_(self._pin_src, non_blocking=True)
        self._load_small(lengths, targets, target_lengths)
        self._mark_uploaded()

    def _wait_uploaded(self) -> None:
        if self._uploaded is not None:
            self._uploaded.synchronize()

    def _mark_uploaded(self) -> None:
        if self._uploaded is None:
            self._uploaded = torch.cuda.Event()
        self._uploaded.record(torch.cuda.current_stream(self.src.device))

    def load_small(self, lengths: torch.Tensor, targets: torch.Tensor, target_lengths: torch.Tensor) -> None:
        """Upload only the step's integer tensors (lengths, targets, target lengths) -- the features come through prefetch()."""
        if targets.shape[0] != self.B or targets.shape[1] > self.L or lengths.numel() != self.B or target_lengths.numel() != self.B:
            raise EecError("GraphedTrainStep: batch / target shape does not fit the captured step")
        if not lengths.is_cuda:
            engine.check_lengths(lengths, self.t_out)
        self._wait_uploaded()
        self._load_small(lengths, targets, target_lengths)
        self._mark_uploaded()

    def _load_small(self, lengths, targets, target_lengths) -> None:
        B, L = self.B, self.L
        if lengths.is_cuda or targets.is_cuda or target_lengths.is_cuda:
            self.lengths.copy_(lengths, non_blocking=True)
            self.targets.fill_(self.pad_token)
            self.targets[:, : targets.shape[1]].copy_(targets, non_blocking=True)
            self.target_lengths.copy_(target_lengths, non_blocking=True)
            return
        # one pinned staging buffer, one H2D copy for the three small integer tensors
        st = self._pin_small
        st[:B].copy_(lengths.reshape(-1))
        st[B:2 * B].copy_(target_lengths.reshape(-1))
        tg = st[2 * B:].view(B, L)
        tg.fill_(self.pad_token)
        tg[:, : targets.shape[1]].copy_(targets)
        self._dev_small.copy_(st, non_blocking=True)
        self.lengths.copy_(self._dev_small[:B])
        self.target_lengths.copy_(self._dev_small[B:2 * B])
        self.targets.copy_(self._dev_small[2 * B:].view(B, L))

    def replay(self) -> torch.Tensor:
        """Re-run the captured step on whatever the static buffers hold; returns the (static) loss tensor."""
        self.graph.replay()
        return self.loss

    # ---- input prefetch: the host->device copy of step i+1's features overlaps step i's kernels ------------------------
    #   step.prefetch(src_pinned)            # any time after the replay of the previous step was enqueued
    #   ...
    #   step.commit_prefetch(); loss = step.replay()
    # The copy lands in a staging buffer on a side stream; commit_prefetch() makes the compute stream wait for it and moves it
    # into the graph's static input with one device-to-device copy (30 MB: ~10 us), so the graph never reads a half-written input.
    def prefetch(self, src: torch.Tensor) -> None:
        if tuple(src.shape) != tuple(self.src.shape):
            raise EecError(f"GraphedTrainStep.prefetch: src shape {tuple(src.shape)} != captured {tuple(self.src.shape)}")
        if getattr(self, "_stage_src", None) is None:
            self._stage_src = torch.empty_like(self.src)
            self._copy_stream = torch.cuda.Stream(device=self.src.device)
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(self.src.device))
        self._copy_stream.wait_event(self._consumed)        # the previous commit has read the staging buffer
        with torch.cuda.stream(self._copy_stream):
            self._stage_src.copy_(src, non_blocking=True)
            self._staged.record(self._copy_stream)

    def commit_prefetch(self) -> None:
        cur = torch.cuda.current_stream(self.src.device)
        cur.wait_event(self._staged)
        self.src.copy_(self._stage_src, non_blocking=True)
        self._consumed.record(cur)

    def __call__(self, src, lengths, targets, target_lengths) -> torch.Tensor:
        self.load_inputs(src, lengths, targets, target_lengths)
        return self.replay()


class GraphedForward:
    """Inference forward (eval mode, no grad) of the first ``n_exits`` exit groups as ONE CUDA graph: the ~35 launches per
    exit group cost more on the host than on the GPU at batch 64, so RTFx is launch-bound when issued eagerly.

        fwd = eec.GraphedForward(model, batch_size=64, t_in=1501)            # all exits
        log_probs = fwd(src, lengths)                                        # (E, B, T', V) fp32, a static buffer

    ``n_exits=e`` captures the truncated encoder (front end + e groups + heads 1..e) that BASELINE's "RTFx per exit" times."""

    def __init__(self, model, batch_size: int, t_in: int, n_exits: Optional[int] = None, n_mels: Optional[int] = None, warmup: int = 2):
        params = list(model.parameters())
        if not params or not params[0].is_cuda:
            raise EecError("GraphedForward: move the model to a CUDA device first (no CPU path)")
        if model.training:
            raise EecError("GraphedForward captures an inference forward: call model.eval() first")
        model._check_supported()
        dev = params[0].device
        self.model = model
        full = model._cfg()
        # a truncated Splitformer keeps the parallel branch of group 0 (early_exit.py:314-356 applies it whatever follows); only the
        # last group's branch disappears with the last group: total_exits tells the engine where "last" is
        self.cfg = engine.Config(n_exits=n_exits or full.n_exits, n_layers=full.n_layers, n_mels=full.n_mels,
                                 splitformer=full.splitformer, precision=full.precision, total_exits=full.n_exits)
        n_mels = n_mels if n_mels is not None else model._features_length
        self.src = torch.zeros(batch_size, n_mels, t_in, dtype=torch.float32, device=dev)
        self.lengths = torch.full((batch_size,), t_in, dtype=torch.int64, device=dev)
        self._pin_len = torch.empty(batch_size, dtype=torch.int64).pin_memory()
        self.t_out = ((t_in - 3) // 2 + 1 - 3) // 2 + 1
        lib = load()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.eec_launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._run()
        self.launches = int(lib.eec_launch_count() - n0)

    def _run(self):
        m = self.model
        with on_device(self.src.device):
            out, _ = engine.model_forward(m._tensor_dict(), m._operands, self.cfg, self.src, self.lengths, False, False)
        return out

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out

    def __call__(self, src: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        if tuple(src.shape) != tuple(self.src.shape) or lengths.numel() != self.src.shape[0]:
            raise EecError(f"GraphedForward: input shape {tuple(src.shape)} != captured {tuple(self.src.shape)}")
        if not lengths.is_cuda:
            engine.check_lengths(lengths, self.t_out)
            self._pin_len.copy_(lengths.reshape(-1))
            lengths = self._pin_len
        self.src.copy_(src, non_blocking=True)
        self.lengths.copy_(lengths, non_blocking=True)
        return self.replay()


class GraphedEarlyExit:
    """Dynamic early-exit inference (``model.forward_early_exit``) as ONE CUDA graph: exit decisions, survivor compaction and the
    per-kernel active-row limits all live on the device, so the captured graph is valid for any mix of exits.

        ee = eec.GraphedEarlyExit(model.eval(), batch_size=64, t_in=1501, threshold=2.5)
        exit_index, tokens, n_tokens, mean_entropy = ee(src, lengths)       # static device buffers; one D2H copy when read
    """

    def __init__(self, model, batch_size: int, t_in: int, threshold: float, n_mels: Optional[int] = None, warmup: int = 2):
        from . import early_exit_infer
        params = list(model.parameters())
        if not params or not params[0].is_cuda:
            raise EecError("GraphedEarlyExit: move the model to a CUDA device first (no CPU path)")
        if model.training:
            raise EecError("GraphedEarlyExit captures an inference forward: call model.eval() first")
        model._check_supported()
        dev = params[0].device
        self.model, self.threshold = model, float(threshold)
        n_mels = n_mels if n_mels is not None else model._features_length
        self.src = torch.zeros(batch_size, n_mels, t_in, dtype=torch.float32, device=dev)
        self.lengths = torch.full((batch_size,), t_in, dtype=torch.int64, device=dev)
        self._pin_len = torch.empty(batch_size, dtype=torch.int64).pin_memory()
        self.t_out = ((t_in - 3) // 2 + 1 - 3) // 2 + 1
        def run():
            with on_device(dev):
                return early_exit_infer.run(model, self.src, self.lengths, self.threshold)
        lib = load()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.eec_launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = run()
        self.launches = int(lib.eec_launch_count() - n0)

    def replay(self):
        self.graph.replay()
        return self.out

    def __call__(self, src: torch.Tensor, lengths: torch.Tensor):
        if tuple(src.shape) != tuple(self.src.shape) or lengths.numel() != self.src.shape[0]:
            raise EecError(f"GraphedEarlyExit: input shape {tuple(src.shape)} != captured {tuple(self.src.shape)}")
        if not lengths.is_cuda:
            engine.check_lengths(lengths, self.t_out)
            self._pin_len.copy_(lengths.reshape(-1))
            lengths = self._pin_len
        self.src.copy_(src, non_blocking=True)
        self.lengths.copy_(lengths, non_blocking=True)
        return self.replay()
