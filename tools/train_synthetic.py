#!/usr/bin/env python
"""The reference's `train.py --decoder_mode ctc` inner loop (train.py:15-92) on the B200 path, end to end, on synthetic audio:

    waveform --eec.Fbank--> (B, 80, T) --eec.Early_conformer / Splitformer (train mode, dropout)--> (E, B, T', 256) log-probs
      --fused multi-exit CTC--> loss --backward--> [NCCL all-reduce per exit group, overlapped] --clip + Noam + AdamW--> next step

one CUDA graph replay per step.  Flag names follow util/conf.py.  LibriSpeech is not available offline, so the batch is synthetic
(noise + tones, random token targets): the numbers to look at are utterances/s and that the loss moves.

    python tools/train_synthetic.py --steps 50
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_synthetic.py --steps 50
"""
import argparse
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "early-exit-transformer_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_type", default="early_conformer", choices=["early_conformer", "splitformer"])
    ap.add_argument("--n_enc_exits", type=int, default=6)
    ap.add_argument("--n_enc_layers_per_exit", type=int, default=2)
    ap.add_argument("--drop_prob", type=float, default=0.1)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=15.0, help="length of every synthetic utterance")
    ap.add_argument("--warmup", type=int, default=25000)
    ap.add_argument("--clip", type=float, default=1.0)
    ap.add_argument("--weight_decay", type=float, default=5e-4)
    ap.add_argument("--adam_eps", type=float, default=1e-9)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()

    import eec
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    cls = eec.Splitformer if args.model_type == "splitformer" else eec.Early_conformer
    model = cls(src_pad_idx=0, n_enc_exits=args.n_enc_exits, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
                d_feed_forward=2048, n_enc_layers=args.n_enc_layers_per_exit, features_length=80, drop_prob=args.drop_prob,
                depthwise_kernel_size=31, device=dev).to(dev).train()
    for mod in model.modules():                     # util/model_utils.py:10-12
        w = getattr(mod, "weight", None)
        if isinstance(w, torch.nn.Parameter) and w.dim() > 1 and not list(mod.children()):
            torch.nn.init.xavier_uniform_(w)
    model.precision = args.precision
    if world > 1:
        eec.distributed.broadcast_parameters(model, 0)
        eec.distributed.OverlappedGradReducer(model)
    opt = eec.FusedNoamAdamW(model, model_size=256, warmup=args.warmup, betas=(0.9, 0.98), eps=args.adam_eps,
                             weight_decay=args.weight_decay, clip=args.clip)

    # synthetic batch: noise + a tone per utterance, targets [<s>, tokens, </s>, pad...] like util/data_loader.py:207-214
    g = torch.Generator().manual_seed(100 + rank)
    B, L = args.batch_size, int(args.seconds * 16000)
    waves = torch.randn(B, L, generator=g) * 0.05
    tt = torch.arange(L) / 16000.0
    for b in range(B):
        waves[b] += 0.2 * torch.sin(2 * math.pi * (200.0 + 40.0 * b) * tt)
    n_samples = torch.randint(L // 2, L + 1, (B,), generator=g)
    n_samples[0] = L
    for b in range(B):
        waves[b, int(n_samples[b]):] = 0
    tl = torch.randint(20, 81, (B,), generator=g)
    targets = torch.full((B, int(tl.max()) + 2), 126, dtype=torch.int64)
    for b in range(B):
        k = int(tl[b])
        targets[b, 0], targets[b, 1 + k] = 1, 2
        targets[b, 1:1 + k] = torch.randint(3, 126, (k,), generator=g)
    tl = tl + 2

    fbank = eec.Fbank()
    feats, frames = fbank(waves.to(dev), n_samples)            # util/data_loader.py:124-125 for the whole batch
    step = eec.GraphedTrainStep(model, B, feats.shape[2], targets.shape[1], optimizer=opt)
    step.load_inputs(feats, frames.cpu(), targets, tl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        feats, frames = fbank(waves.to(dev, non_blocking=True), n_samples)    # features recomputed every step, as a loader would
        step.src.copy_(feats, non_blocking=True)
        loss = step.replay()
        if rank == 0 and (i % 10 == 0 or i == args.steps - 1):
            print(f"step {i:4d}  loss {float(loss):9.4f}", flush=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"{args.steps} steps, {world} GPU(s): {world * B * args.steps / dt:.0f} utt/s ({dt / args.steps * 1e3:.2f} ms/step incl. fbank + H2D), "
              f"rate {opt.rate():.3e}")
    if world > 1:
        step = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
