#!/usr/bin/env python
"""Data-parallel exchange check on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Every rank runs the same weights on its own synthetic batch and computes the averaged gradients three ways:
  (a) eager backward, then ONE all-reduce of the flat gradient buffer (eec.distributed.all_reduce_gradients);
  (b) eager backward with eec.distributed.OverlappedGradReducer (per-exit-group all-reduce on a side stream during backward);
  (c) the same as (b) captured into the training CUDA graph (GraphedTrainStep) and replayed twice.
(b) and (c) must equal (a) up to the summation order of the atomics inside backward.  Prints PASS / FAIL on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "early-exit-transformer_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import bench
    import eec
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bn, layers = 8, 1
    bench.B = Bn
    model = bench.build_model(layers, "bf16", dev).train()
    eec.distributed.broadcast_parameters(model, 0)
    src, lengths, targets, tl = bench.synthetic_batch(Bn, 99 + rank)
    src_d, tg_d, tl_d = src.to(dev), targets.to(dev), tl.to(dev)

    def eager():
        model.zero_grad(set_to_none=True)
        out = model(src_d, lengths)
        loss = eec.multi_exit_ctc_loss(out, tg_d, tl_d)
        loss.backward()
        return loss.detach().clone()     # (do not keep the autograd graph alive: its AccumulateGrad nodes are bound to this stream)

    la = eager()
    local_flat = model._flat_grad.clone()
    eec.distributed.all_reduce_gradients(model)
    ref = model._flat_grad.clone()
    changed = float((ref - local_flat).abs().max())          # ranks have different batches: averaging must change the buffer

    red = eec.distributed.OverlappedGradReducer(model)
    lb = eager()
    torch.cuda.synchronize()
    gb = model._flat_grad.clone()
    calls_eager = red.calls

    step = eec.GraphedTrainStep(model, Bn, src.shape[2], targets.shape[1])
    step(src, lengths, targets, tl)
    step(src, lengths, targets, tl)
    torch.cuda.synchronize()
    gc = model._flat_grad.clone()

    scale = float(ref.abs().max())
    eb, ec = float((gb - ref).abs().max()) / scale, float((gc - ref).abs().max()) / scale
    t = torch.tensor([eb, ec, changed / scale], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    same = ref.clone()
    dist.all_reduce(same, op=dist.ReduceOp.MAX)              # every rank must hold the same averaged buffer
    agree = float((same - ref).abs().max()) / scale
    ok = t[0] < 2e-3 and t[1] < 2e-3 and t[2] > 1e-3 and agree < 1e-6
    if rank == 0:
        print(f"dp_check world={world}: |overlapped-eager - flat|/max = {float(t[0]):.2e}, |overlapped-graph - flat|/max = {float(t[1]):.2e}, "
              f"averaging changed the local buffer by {float(t[2]):.2e}, rank agreement {agree:.1e}, slices per backward {calls_eager}, "
              f"loss {float(la):.4f}/{float(lb):.4f}, launches in graph {step.launches_per_step}: {'PASS' if ok else 'FAIL'}", flush=True)
    # (a captured graph that holds NCCL kernels must be released before the communicator goes away)
    del step
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
