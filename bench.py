#!/usr/bin/env python
"""bench.py -- headline benchmark of the early-exit conformer hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # our arm (libeec.so kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): early_conformer CTC training step -- forward, summed 6-exit CTC
loss, backward -- bf16, batch 64 x 15 s (T_in=1501 -> T'=374), 12 layers / 6 exits, d_model 256,
synthetic fbank + random-init weights.  N>1: batch-sharded data parallel, 64 utterances
per GPU (weak scaling), one NCCL all-reduce of the flat fp32 gradient buffer per step.
Prints ONE JSON line (rank 0).  A "step" = one pass of the hot path over one batch.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "early-exit-transformer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

B, T_IN, N_MELS = 64, 1501, 80
N_EXITS = 6
FRAME_S = 0.010  # hop 160 @ 16 kHz (util/conf.py:335-341)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sust": p["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sust": 1400.0, "src": "fallback"}


def t_out(t_in):
    return ((t_in - 3) // 2 + 1 - 3) // 2 + 1


def model_flops_fwd(n_frames, t_enc, layers):
    per_layer = 2 * 2097152 + 393216 + 131072 + 4 * t_enc * 256 + 262144 + 131072
    return n_frames * (layers * per_layer + N_EXITS * 131072) + n_frames * 2 * (256 * 768 + 2 * 256 * 240)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            text = ""
        sm, mx, reasons = [], [], set()
        for line in text.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_state_dict(layers_per_exit, splitformer=False):
    """Random-init weights of the named architecture: the host-side mirror of the reference class is built on
    the CPU under torch.manual_seed(0) (same default init as the reference's ctor), every >= 2-D `.weight` of a leaf module gets
    Xavier-uniform like util/model_utils.py:10-12 (model.apply(initialize_weights), train.py:229-230), and 1-D parameters /
    BatchNorm buffers are randomised so that no scale or shift is a no-op.  No CUDA work, nothing from oracle/."""
    import eec
    torch.manual_seed(0)
    cls = eec.Splitformer if splitformer else eec.Early_conformer
    m = cls(src_pad_idx=0, n_enc_exits=N_EXITS, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
            max_len=2000, d_feed_forward=2048, n_enc_layers=layers_per_exit, features_length=N_MELS,
            drop_prob=0.0, depthwise_kernel_size=31, device=torch.device("cpu"))
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for mod in m.modules():
            w = getattr(mod, "weight", None)
            if isinstance(w, torch.nn.Parameter) and w.dim() > 1 and not list(mod.children()):
                torch.nn.init.xavier_uniform_(w, generator=g)
        for n, prm in m.named_parameters():
            if prm.dim() == 1:
                base = 1.0 if (n.endswith("norm.weight") or ".sequential.0.weight" in n and "ffn" in n or n.endswith("sequential.3.weight")) else 0.0
                prm.copy_(base + 0.05 * (2 * torch.rand(prm.shape, generator=g) - 1))
        for n, buf in m.named_buffers():
            if n.endswith("running_mean"):
                buf.copy_(0.1 * (2 * torch.rand(buf.shape, generator=g) - 1))
            elif n.endswith("running_var"):
                buf.copy_(1.0 + 0.2 * torch.rand(buf.shape, generator=g))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def build_model(layers_per_exit, precision, device, splitformer=False):
    import eec
    cls = eec.Splitformer if splitformer else eec.Early_conformer
    m = cls(src_pad_idx=0, n_enc_exits=N_EXITS, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
            max_len=2000, d_feed_forward=2048, n_enc_layers=layers_per_exit, features_length=N_MELS,
            drop_prob=0.0, depthwise_kernel_size=31, device=device)
    m.load_state_dict(synthetic_state_dict(layers_per_exit, splitformer), strict=True)
    m = m.to(device)
    m.precision = precision
    return m


def synthetic_batch(n, seed):
    """randn fbank (n, 80, 1501), lengths ~ U[750, 1501] with lengths[0] = 1501, padding zeroed
    (util/data_loader.py:21-26); targets [<s>=1, tokens in 3..125, </s>=2, pad 126...] with 20..80 tokens (:207-214)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, N_MELS, T_IN, generator=g)
    lengths = torch.randint(T_IN // 2, T_IN + 1, (n,), generator=g)
    lengths[0] = T_IN
    for b in range(n):
        src[b, :, int(lengths[b]):] = 0.0
    tl = torch.randint(20, 81, (n,), generator=g)
    tg = torch.full((n, int(tl.max()) + 2), 126, dtype=torch.int64)
    for b in range(n):
        k = int(tl[b])
        tg[b, 0] = 1
        tg[b, 1:1 + k] = torch.randint(3, 126, (k,), generator=g)
        tg[b, 1 + k] = 2
    return src, lengths.to(torch.int64), tg, (tl + 2).to(torch.int64)


def synthetic(rank):
    return synthetic_batch(B, 1234 + rank)


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_reference():
    """The UNMODIFIED reference (models/model/early_exit.py Early_conformer + util/noam_opt.py NoamOpt) from baseline/_ref
    (byte-for-byte copies made by baseline/install_ref.py; git-ignored, shipped to the GPU box), or None when that directory is
    absent -- the callers then fall back to the oracle port and say kind = "port"."""
    if not os.path.isdir(os.path.join(REF_DIR, "models")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        from models.model.early_exit import Early_conformer   # noqa: E402  (the reference)
        from util.noam_opt import NoamOpt                      # noqa: E402
    except Exception:   # noqa: BLE001
        return None
    return Early_conformer, NoamOpt


def reference_model(ref, layers, device):
    """the reference class with its own ctor call (train.py:166-178) and the benchmark's random-init weights"""
    m = ref[0](src_pad_idx=0, n_enc_exits=N_EXITS, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
               d_feed_forward=2048, n_enc_layers=layers, features_length=N_MELS, drop_prob=0.0, depthwise_kernel_size=31,
               device=device)
    m.load_state_dict(synthetic_state_dict(layers), strict=True)
    return m.to(device)


class ReferenceTrainer:
    """The reference's own training step, statement for statement (train.py:53-70 with --decoder_mode ctc; loss and optimiser as
    built at train.py:258-262): model(src, lengths) -> six nn.CTCLoss calls summed -> zero_grad -> backward -> clip_grad_norm_ ->
    NoamOpt(AdamW).step().  Nothing of this repo's kernels, engine or model mirror is on this path."""

    def __init__(self, ref, layers, device, autocast=False):
        self.model = reference_model(ref, layers, device).train()
        self.ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)
        self.opt = ref[1](256, 25000, torch.optim.AdamW(params=self.model.parameters(), lr=0, betas=(0.9, 0.98), eps=1e-9,
                                                        weight_decay=5e-4))
        self.device, self.autocast = device, autocast

    def step(self, src, lengths, targets, tl):
        with torch.autocast(torch.device(self.device).type, dtype=torch.bfloat16, enabled=self.autocast):
            encoder = self.model(src, lengths)
        in_len = torch.full(size=(encoder.size(1),), fill_value=encoder.size(2), dtype=torch.long)
        loss = 0
        for enc in encoder:
            loss += self.ctc(enc.float().permute(1, 0, 2), targets, in_len, tl)
        self.model.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)
        with open(os.devnull, "w") as null:     # (NoamOpt.step prints "RATE: ..." every call, util/noam_opt.py:33)
            import contextlib
            with contextlib.redirect_stdout(null):
                self.opt.step()
        return loss


def workload_config(layers):
    """`config` of the JSON line: ONE definition for both arms (ours and --impl reference), so the driver sees the same workload."""
    T = t_out(T_IN)
    return {"workload": f"early_conformer CTC training step (fwd + summed {N_EXITS}-exit CTC + bwd + clip_grad_norm + Noam/AdamW update), "
                        f"{N_EXITS} exits x {layers} layers, d_model 256, 15 s utterances (T_in {T_IN} -> T' {T}), batch {B} per model "
                        f"call, dropout 0 (parity configuration), synthetic fbank + random-init weights",
            "batch_per_gpu": B, "t_in": T_IN, "exits_x_layers": f"{N_EXITS}x{layers}",
            "l2": "per-step working set ~8 GB >> 126 MB L2 (no flush needed)"}


class TrainLeg:
    """One complete training configuration on this rank: model, fused optimiser, [overlapped gradient reducer,] the step's CUDA graph."""

    def __init__(self, args, layers, dev, world, rank, drop_p=0.0):
        import torch.distributed as dist
        import eec
        self.args, self.world, self.dist, self.eec = args, world, dist, eec
        self.model = model = build_model(layers, args.precision, dev).train()
        model.dropout = drop_p
        self.src, self.lengths, self.targets, self.tl = synthetic(rank)
        self.tg_dev, self.tl_dev = self.targets.to(dev), self.tl.to(dev)
        # optimiser: clip_grad_norm_ + Noam + AdamW as three launches over flat buffers, reference hyper-parameters
        self.opt = None if args.no_opt else eec.FusedNoamAdamW(model, model_size=256, warmup=25000, betas=(0.9, 0.98), eps=1e-9,
                                                               weight_decay=5e-4, clip=1.0)
        # data parallel: the gradient all-reduce runs DURING backward, one exit group at a time on a side stream
        # (eec.distributed.OverlappedGradReducer); the collectives are stream-ordered, so they are captured into the step's graph
        self.overlap = world > 1 and not args.no_overlap
        if world > 1:
            eec.distributed.broadcast_parameters(model, 0)
            if self.opt is not None:
                self.opt.refresh_shadow()
        if self.overlap:
            eec.distributed.OverlappedGradReducer(model, grad_dtype=args.grad_dtype)
        self.in_graph = world == 1 or self.overlap      # is the whole step (incl. exchange + update) one graph?
        self.graphed = None
        self.parity = None

    def capture(self, dev):
        # the whole step (forward, 6-exit CTC, backward, [overlapped all-reduce,] clip + Noam + AdamW) is ONE CUDA graph.
        # With --no-overlap the single flat-buffer all-reduce sits between the graph and an eager update.
        if not self.args.no_graph:
            self.graphed = self.eec.GraphedTrainStep(self.model, B, T_IN, self.targets.shape[1], optimizer=self.opt if self.in_graph else None)
            self.graphed.load_inputs(self.src.to(dev), self.lengths, self.tg_dev, self.tl_dev)
            self.src_dev = self.graphed.src   # timed steps run on inputs already resident in the graph's static buffer
        else:
            self.src_dev = self.src.to(dev)

    def step(self, x=None):
        g, model, opt = self.graphed, self.model, self.opt
        if g is not None:
            if x is not None and x is not g.src:
                g.src.copy_(x, non_blocking=True)
            loss = g.replay()
        else:
            out = model(self.src_dev if x is None else x, self.lengths)
            loss = self.eec.multi_exit_ctc_loss(out, self.tg_dev, self.tl_dev)
            model.zero_grad(set_to_none=True)
            loss.backward()
            if opt is not None and self.in_graph:
                opt.step()
        if not self.in_graph:
            self.dist.all_reduce(model._flat_grad, op=self.dist.ReduceOp.AVG)
            if opt is not None:
                opt.step()
        return loss

    @property
    def launches(self):
        return self.graphed.launches_per_step if self.graphed is not None else None

    def close(self):
        self.graphed = None


def run_ours(args):
    import torch.distributed as dist
    import eec
    from eec import lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} != --gpus {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eec.load()
    pk = peaks()
    layers = args.layers_per_exit
    T = t_out(T_IN)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # auxiliary legs never take the headline down with them: a failure is reported in place of the leg's numbers
    def guarded(fn, *a, **k):
        try:
            return fn(*a, **k)
        except Exception as e:   # noqa: BLE001
            return {"failed": f"{type(e).__name__}: {e}"[:300]}

    leg = TrainLeg(args, layers, dev, world, rank, drop_p=args.train_drop_prob)
    model, lengths = leg.model, leg.lengths
    # parity of the benchmarked step, in the run that times it (rank 0; initial weights; the reference's CPU forward on the same batch)
    parity = None
    if rank == 0 and not args.skip_cpu and not args.skip_parity and not args.profile:
        parity = guarded(parity_leg, model, layers, leg.src, lengths, leg.targets, leg.tl, dev)
    leg.capture(dev)
    src_pin = leg.src.pin_memory()
    graphed = leg.graphed
    warm = max(args.warmup, 3)
    for _ in range(warm):
        leg.step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.load().eec_launch_count()
    if args.profile:
        torch.cuda.profiler.start()
    ms = timed(leg.step, args.steps)
    if args.profile:
        torch.cuda.profiler.stop()
    launches = leg.launches if graphed is not None else (L.load().eec_launch_count() - launches0) // args.steps
    clocks = sampler.stop() if rank == 0 else None

    # end to end: every step uploads one batch of features from pinned host memory and reads its loss back.  With the graph the
    # upload of step i+1 is issued right after step i's replay (GraphedTrainStep.prefetch: side stream -> staging buffer) and is
    # committed to the graph's static input at the start of step i+1, so it overlaps step i's kernels like a prefetching loader.
    # lengths / targets / target lengths (5.8 KB) are uploaded with every step as well (GraphedTrainStep.load_small).
    small_bytes = (leg.lengths.numel() + leg.targets.numel() + leg.tl.numel()) * 8

    def e2e_step():
        if graphed is not None:
            graphed.commit_prefetch()
            graphed.load_small(leg.lengths, leg.targets, leg.tl)
            loss = leg.step()
            graphed.prefetch(src_pin)
            return float(loss.item())
        x = src_pin.to(dev, non_blocking=True)
        return float(leg.step(x).item())

    if args.profile:
        ms_e2e = float("nan")
    else:
        if graphed is not None:
            graphed.prefetch(src_pin)
        e2e_step()
        ms_e2e = timed(e2e_step, args.steps)

    # DP step kernel list (torch.profiler / CUPTI on rank 0; ncu cannot follow a multi-rank job): written next to the JSON line
    if args.kernel_trace and graphed is not None:
        guarded(kernel_trace, leg.step, args.kernel_trace, rank)

    # inference RTFx per exit (BASELINE metric part (i)): forward truncated after exit e, bf16, eval
    rtfx = ee_leg = fb_leg = eager_leg = hbm_leg = f32_leg = beam_leg = aed = None
    src_dev = leg.src_dev
    audio_s = float(lengths.sum()) * FRAME_S
    if rank == 0 and not args.skip_rtfx and not args.profile:
        rtfx = rtfx_per_exit(model, src_dev, lengths, audio_s, use_graph=not args.no_graph)
        model.train()
        ee_leg = guarded(early_exit_leg, layers, args.precision, dev, src_dev, lengths, audio_s) if not args.no_graph else None
        fb_leg = guarded(fbank_leg, dev, with_cpu=(world == 1 and not args.skip_cpu))
        hbm_leg = guarded(roofline_hbm_leg, dev, pk)
        if world == 1 and not args.no_graph and args.precision == "bf16":
            f32_leg = guarded(fp32_leg, layers, dev, src_dev, lengths, leg.tg_dev, leg.tl_dev, audio_s, leg.targets.shape[1])
        if world == 1 and not args.skip_cpu:
            eager_leg = guarded(torch_eager_leg, layers, dev, src_dev, lengths, leg.tg_dev, leg.tl_dev, audio_s)
        beam_leg = guarded(beam_search_leg, model, src_dev, lengths) if not args.no_graph else None

    # the same step at the reference's DEFAULT --drop_prob 0.1 (util/conf.py:283-291): fused counter-based dropout at all
    # seven sites per layer + after the positional encoding, masks regenerated in backward (nothing stored)
    drop_leg = None
    if args.drop_prob > 0 and not args.profile and graphed is not None:
        model.train()
        model.dropout = args.drop_prob
        g2 = eec.GraphedTrainStep(model, B, T_IN, leg.targets.shape[1], optimizer=leg.opt if leg.in_graph else None)
        g2.load_inputs(src_dev, lengths, leg.tg_dev, leg.tl_dev)

        def dstep():
            loss = g2.replay()
            if not leg.in_graph:
                dist.all_reduce(model._flat_grad, op=dist.ReduceOp.AVG)
                if leg.opt is not None:
                    leg.opt.step()
            return loss
        for _ in range(3):
            dstep()
        ms_d = timed(dstep, args.steps)
        drop_leg = {"drop_prob": args.drop_prob, "ms_per_step": round(ms_d / args.steps, 3),
                    "value": round(world * B / (ms_d / args.steps / 1e3), 2), "unit": "utt/s", "gpu_launches": g2.launches_per_step}
        model.dropout = 0.0
        del g2

    # BASELINE configs[2]: the 18-layer (6 exits x 3 layers) model, data parallel at N = 2 / 4 / 8 (and at N = 1 for the efficiency's
    # denominator): the same complete step, gradient slices of 31 MB per exit group (188 MB per step)
    deep_leg = None
    if layers != 3 and not args.skip_deep and not args.profile and graphed is not None:
        leg.close()
        graphed = None
        del leg
        torch.cuda.empty_cache()

        def deep():
            dl = TrainLeg(args, 3, dev, world, rank)
            dl.capture(dev)
            for _ in range(3):
                dl.step()
            ms_deep = timed(dl.step, args.steps) / args.steps
            fl = 3.0 * model_flops_fwd(B * T, T, N_EXITS * 3)
            res = {"config": "BASELINE configs[2]: early_conformer 6 exits x 3 layers (18 layers), batch 64 per GPU, data parallel "
                             f"dp{world}", "ms_per_step": round(ms_deep, 3), "value": round(world * B / (ms_deep / 1e3), 2), "unit": "utt/s",
                   "n_gpus": world, "gpu_launches": dl.launches, "grad_bytes_all_reduced_per_step": 46977536 * (2 if args.grad_dtype == "bf16" else 4) if world > 1 else 0,
                   "step_tensor_frac_of_sustained": round(fl / (ms_deep / 1e3) / 1e12 / pk["tf_sust"], 4)}
            dl.close()
            return res
        deep_leg = guarded(deep)
        if world > 1:
            barrier()

    if rank == 0 and world == 1 and not args.skip_aed and not args.profile and not args.skip_rtfx:
        if graphed is not None:
            leg.close()
            graphed = None
        torch.cuda.empty_cache()
        aed = guarded(aed_leg, dev)
    roof = cpu = None
    if rank == 0:
        roof = None if args.profile else roofline_dominant(dev, pk)
        if world == 1 and not args.skip_cpu and not args.profile:
            cpu = cpu_baseline(layers, sample_b=args.cpu_sample)

    if rank == 0:
        per_step = ms / args.steps
        value = world * B / (per_step / 1e3)
        flops = 3.0 * model_flops_fwd(B * T, T, N_EXITS * layers)
        overlap = world > 1 and not args.no_overlap
        cfg = workload_config(layers)
        line = {
            "metric": "train_utts_per_sec", "value": round(value, 2), "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": round(per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": cfg,
            "details": {"global_batch": world * B, "parallelism": f"dp{world}",
                        "grad_all_reduce": (f"NCCL {args.grad_dtype}, per exit group, overlapped with backward inside the graph" if overlap
                                            else "NCCL fp32 flat buffer after backward") if world > 1 else "n/a",
                        "optimizer": "excluded" if args.no_opt else "clip_grad_norm + Noam + AdamW update included (fused, flat buffers)",
                        "dropout": args.train_drop_prob,
                        "launch": "eager (ctypes launches)" if args.no_graph else "one CUDA graph replay per step",
                        "step_tflops_algorithmic": round(flops / 1e12, 3),
                        "step_tensor_frac_of_sustained": round(flops / (per_step / 1e3) / 1e12 / pk["tf_sust"], 4),
                        "peaks": pk["src"]},
            "e2e": {"value": round(world * B / (ms_e2e / args.steps / 1e3), 2), "unit": "utt/s",
                    "h2d_bytes_per_step": src_pin.numel() * 4 + small_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "parity_vs_reference": parity,
        }
        for key, val in (("rtfx_per_exit", rtfx), ("train_with_dropout", drop_leg), ("dp_18_layers", deep_leg), ("aed_mode", aed),
                         ("ctc_beam_search", beam_leg),
                         ("early_exit_inference", ee_leg), ("fbank_frontend", fb_leg), ("roofline_hbm_kernel", hbm_leg),
                         ("fp32_mode", f32_leg), ("torch_eager_same_gpu", eager_leg)):
            if val is not None:
                line[key] = val
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        graphed = None          # (graphs that hold captured NCCL kernels are released before the communicator)
        torch.cuda.synchronize()
        dist.barrier()
        # Hard exit instead of dist.destroy_process_group(): with CUDA graphs that captured NCCL kernels still alive in several legs, the
        # communicator teardown hung intermittently AFTER the JSON line was out (2 of 9 multi-GPU runs in round 2: the line was complete,
        # the processes never left).  The result is printed and flushed, every rank has passed the barrier; the driver reclaims the rest.
        _JSON_OUT.flush()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def kernel_trace(step, path, rank):
    """Kernel list of ONE training step on rank 0 through torch.profiler (CUPTI activity records: names + device durations, including the
    ncclDevKernel_* launches captured in the step's graph) -- ncu replays kernels and cannot profile a multi-rank job."""
    from torch.profiler import ProfilerActivity, profile
    if rank != 0:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        return
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type is not None and str(ev.device_type).endswith("CUDA") and ev.device_time > 0:
            r = rows.setdefault(ev.name, [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time
    tot = sum(v[1] for v in rows.values())
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    with open(path, "w") as f:
        f.write("# torch.profiler (CUPTI) kernel list of 3 replays of the training-step graph, rank 0; device time per launch\n")
        f.write(f"# total kernel time {tot / 3 / 1e3:.3f} ms per step\n")
        f.write("share%  total_us_per_step  launches_per_step  avg_us  kernel\n")
        for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * us / tot:6.2f}  {us / 3:10.1f}  {n / 3:8.1f}  {us / n:9.2f}  {name[:160]}\n")


def parity_leg(model, layers, src, lengths, targets, tl, dev):
    """The HEADLINE step checked in the run that times it: train-mode log-probabilities of every exit and the summed 6-exit CTC loss of the
    benchmarked model on the benchmarked batch (B = 64, initial weights), from the CUDA path and from the reference's CPU path (baseline/_ref,
    unmodified, fp32; oracle port when absent) -- relative errors per exit (maxabs(a-b)/maxabs(b), the north star's definition) and of the loss."""
    import eec
    bufs = {n: b.detach().clone() for n, b in model.named_buffers()}      # a train-mode forward moves the BatchNorm running stats:
    with torch.no_grad():
        out = model(src.to(dev), lengths)
        loss = float(eec.multi_exit_ctc_loss(out, targets, tl))
        for n, b in model.named_buffers():                                  # ... put them back, the timed steps start from the same state
            b.copy_(bufs[n])
    out = out.cpu()
    t0 = time.perf_counter()
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    ref = load_reference()
    with torch.no_grad():
        if ref is not None:
            kind = "reference"
            r = reference_model(ref, layers, torch.device("cpu")).train()(src, lengths)
        else:
            from oracle import conformer_oracle as O
            kind = "port"
            r = O.early_conformer_forward(synthetic_state_dict(layers), src, lengths, training=True, bn_out={})
        in_len = torch.full((r.shape[1],), r.shape[2], dtype=torch.long)
        ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)
        rloss = float(sum(ctc(enc.permute(1, 0, 2), targets, in_len, tl) for enc in r))
    per_exit = [round(float((out[e] - r[e]).abs().max() / r[e].abs().max()), 6) for e in range(r.shape[0])]
    tol = 2e-2 if model.precision == "bf16" else 1e-3
    return {"against": f"the reference's CPU forward (kind {kind}, fp32, train-mode BatchNorm) on the same {src.shape[0]}-utterance batch and weights",
            "kind": kind, "logprob_rel_err_per_exit": per_exit, "loss": round(loss, 5), "loss_reference": round(rloss, 5),
            "loss_rel_err": round(abs(loss - rloss) / abs(rloss), 7), "tolerance": tol,
            "ok": bool(max(per_exit) < tol and abs(loss - rloss) / abs(rloss) < tol), "cpu_seconds": round(time.perf_counter() - t0, 1)}


def _time_loop(run, n, warm=3):
    for _ in range(warm):
        run()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n):
        run()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n


def fp32_leg(layers, dev, src_dev, lengths, tg_dev, tl_dev, audio_s, targets_w):
    """model.precision = "fp32" (the mirror's DEFAULT, what INTEGRATION.md's two-line switch gives a maintainer before opting into bf16):
    CUDA-core FFMA GEMMs and attention, reference-accurate (greedy tokens bit-exact).  Training step and six-exit forward, one graph each."""
    import eec
    m = build_model(layers, "fp32", dev).train()
    opt = eec.FusedNoamAdamW(m, model_size=256, warmup=25000, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-4, clip=1.0)
    g = eec.GraphedTrainStep(m, B, T_IN, targets_w, optimizer=opt)
    g.load_inputs(src_dev, lengths, tg_dev, tl_dev)
    ms_train = _time_loop(g.replay, 3, warm=2)
    del g, opt
    m.eval()
    with torch.no_grad():
        fwd = eec.GraphedForward(m, B, T_IN)
        fwd(src_dev, lengths)
        ms_fwd = _time_loop(fwd.replay, 5, warm=2)
    del fwd, m
    torch.cuda.empty_cache()
    return {"precision": "fp32 (FFMA kernels, no tensor cores)", "train_ms_per_step": round(ms_train, 2), "train_utt_per_s": round(B / (ms_train / 1e3), 1),
            "inference_all_exits_ms": round(ms_fwd, 2), "inference_rtfx": round(audio_s / (ms_fwd / 1e3), 1)}


def aed_leg(dev, steps=3):
    """BASELINE configs[4]: early_conformer AED mode -- full_conformer = the 12-layer / 6-exit encoder with per-exit CTC heads plus six
    6-layer attention decoders (88.8 M parameters), batch 64, long utterances at the model's max_len (T_in 7999 -> T' 1999 encoder
    frames), targets of 150..400 tokens.  One training step: forward, 0.7 CE + 0.3 CTC over all exits (train.py:36-51), backward (no optimiser:
    the fused one covers the CTC models).  Encoder AND decoder stacks on the sm_100a kernels; the reference's own module in eager PyTorch
    (bf16 autocast) beside it when it fits."""
    import eec
    Bn, T_in = 64, 7999
    g = torch.Generator().manual_seed(77)
    lengths = torch.randint(T_in // 2, T_in + 1, (Bn,), generator=g)
    lengths[0] = T_in
    src = torch.randn(Bn, N_MELS, T_in, generator=g)
    for b in range(Bn):
        src[b, :, int(lengths[b]):] = 0.0
    tl = torch.randint(150, 401, (Bn,), generator=g)
    tg = torch.full((Bn, int(tl.max()) + 2), 126, dtype=torch.int64)
    for b in range(Bn):
        k = int(tl[b])
        tg[b, 0] = 1
        tg[b, 1:1 + k] = torch.randint(3, 126, (k,), generator=g)
        tg[b, 1 + k] = 2
    tl = tl + 2
    kw = dict(trg_pad_idx=126, n_enc_exits=N_EXITS, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000, d_feed_forward=2048,
              n_enc_layers=2, n_dec_layers=6, features_length=N_MELS, drop_prob=0.0, depthwise_kernel_size=31)
    torch.manual_seed(0)
    m = eec.full_conformer(device=dev, **kw).to(dev).train()
    m.precision = "bf16"
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    src_d, tg_d, tl_d = src.to(dev), tg.to(dev), tl.to(dev)
    trg, trg_expect = tg_d[:, :-1].contiguous(), tg_d[:, 1:].contiguous()

    def step():
        dec, enc = m(src_d, lengths, trg)
        loss = 0.7 * eec.multi_exit_cross_entropy(dec, trg_expect) + 0.3 * eec.multi_exit_ctc_loss(enc, tg_d, tl_d)
        m.zero_grad(set_to_none=True)
        loss.backward()
        return loss
    n0 = eec.load().eec_launch_count()
    loss0 = float(step())
    launches = int(eec.load().eec_launch_count() - n0)
    ms = _time_loop(step, steps, warm=1)
    res = {"config": "BASELINE configs[4]: full_conformer (AED), 6 exits x 2 encoder layers + 6 x 6 decoder layers, batch 64, T_in 7999 -> T' 1999 "
                     f"(max_len 2000), target width {tg.shape[1]}; fwd + (0.7 CE + 0.3 CTC) + bwd, bf16", "ms_per_step": round(ms, 2),
           "utt_per_s": round(Bn / (ms / 1e3), 1), "loss_first_step": round(loss0, 4), "gpu_launches": launches,
           "parameters": sum(p.numel() for p in m.parameters())}
    del m
    torch.cuda.empty_cache()
    ref = load_reference()
    if ref is not None:
        try:
            if REF_DIR not in sys.path:
                sys.path.insert(0, REF_DIR)
            from models.model.early_exit import full_conformer as RefFC   # noqa: E402
            torch.manual_seed(0)
            r = RefFC(device=dev, **kw).to(dev).train()
            r.load_state_dict(sd, strict=True)
            ctc, ce = torch.nn.CTCLoss(blank=0, zero_infinity=True), torch.nn.CrossEntropyLoss()

            def rstep():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    dec, enc = r(src_d, lengths, trg)
                in_len = torch.full((Bn,), enc.size(2), dtype=torch.long)
                lc = sum(ctc(e_.float().permute(1, 0, 2), tg_d, in_len, tl_d) for e_ in enc)
                lce = sum(ce(d_.float().permute(0, 2, 1), trg_expect) for d_ in dec)
                loss = 0.7 * lce + 0.3 * lc
                r.zero_grad()
                loss.backward()
                return loss
            rl = float(rstep())
            res["reference_eager_bf16_autocast_ms_per_step"] = round(_time_loop(rstep, 2, warm=0), 2)
            res["reference_loss_first_step"] = round(rl, 4)
            res["loss_rel_err_vs_reference_autocast"] = round(abs(loss0 - rl) / abs(rl), 5)
            del r
        except Exception as e:   # noqa: BLE001
            res["reference_eager"] = f"failed: {type(e).__name__}: {e}"[:200]
        torch.cuda.empty_cache()
    return res


def rtfx_per_exit(model, src_dev, lengths, audio_s, use_graph=True):
    """Inference RTFx for a forward truncated after exit e (e = 1..6): front end + e exit groups + heads 1..e, bf16, eval;
    each truncated forward is one CUDA graph replay (eager launches with --no-graph)."""
    import eec
    from eec import engine
    model.eval()
    res = []
    P, W = model._tensor_dict(), model._operands
    full = model._cfg()
    with torch.no_grad():
        for e in range(1, N_EXITS + 1):
            if use_graph:
                fwd = eec.GraphedForward(model, src_dev.shape[0], src_dev.shape[2], n_exits=e)
                fwd(src_dev, lengths)
                run = fwd.replay
            else:
                cfg = engine.Config(n_exits=e, n_layers=full.n_layers, n_mels=full.n_mels, precision=full.precision)
                run = lambda: engine.model_forward(P, W, cfg, src_dev, lengths, False, False)   # noqa: E731
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(5):
                run()
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 5
            res.append({"exit": e, "ms": round(ms, 3), "rtfx": round(audio_s / (ms / 1e3), 1)})
    return res


def beam_search_leg(model, src_dev, lengths):
    """what `inference.py --decoder_mode ctc` prints -- the beam-10 CTC prefix beam search of ALL six exits of the batch (384
    emission matrices of 374 x 256) as ONE kernel launch behind the six-exit forward; torchaudio's cuda_ctc_decoder (the reference's
    library, util/beam_infer.py:100-110) beside it where it runs (it faults on sm_100 for beam_size >= 7, so it is timed at beam 5)."""
    import eec
    model.eval()
    with torch.no_grad():
        out = eec.GraphedForward(model, src_dev.shape[0], src_dev.shape[2])(src_dev, lengths).clone()
    model.train()
    vocab = [str(i) for i in range(out.shape[-1])]
    res = {"emissions": list(out.shape)}
    for beam in (10, 5):
        dec = eec.cuda_ctc_decoder(vocab, nbest=1, beam_size=beam, blank_skip_threshold=0.95)
        ms = _time_loop(lambda: dec.search(out), 5, warm=2)
        res[f"all_exits_beam{beam}_ms"] = round(ms, 3)
    try:
        import subprocess
        code = ("import sys,time,torch\nfrom torchaudio.models.decoder import cuda_ctc_decoder\n"
                "g=torch.Generator().manual_seed(0)\nlp=torch.log_softmax(torch.randn(64,374,256,generator=g),-1).cuda()\n"
                "lens=torch.full((64,),374,dtype=torch.int32).cuda()\nd=cuda_ctc_decoder([str(i) for i in range(256)],nbest=1,beam_size=5,blank_skip_threshold=0.95)\n"
                "d(lp,lens);torch.cuda.synchronize();t=time.perf_counter()\nfor _ in range(3): d(lp,lens)\ntorch.cuda.synchronize();print((time.perf_counter()-t)/3*1e3)")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        if r.returncode == 0:
            res["torchaudio_cuda_ctc_decoder_beam5_ms_per_exit"] = round(float(r.stdout.strip().splitlines()[-1]), 2)
            res["torchaudio_all_exits_beam5_ms"] = round(6 * float(r.stdout.strip().splitlines()[-1]), 2)
        else:
            res["torchaudio_cuda_ctc_decoder"] = "failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:160]
    except Exception as e:   # noqa: BLE001
        res["torchaudio_cuda_ctc_decoder"] = f"failed: {type(e).__name__}"
    return res


def early_exit_leg(layers, precision, dev, src_dev, lengths, audio_s):
    """BASELINE configs[3]: Splitformer CTC inference with dynamic early exit and on-device batch compaction, one CUDA graph.
    Random-init weights have no meaningful confidence, so the threshold is set from the measured entropies such that about half
    of the utterances leave within the first three exits: the leg measures the MECHANISM (exit decision, compaction, later
    layers skipping finished rows without a host sync), not an accuracy/latency trade-off."""
    import eec
    m = build_model(layers, precision, dev, splitformer=True).eval()
    with torch.no_grad():
        _, _, _, H = m.forward_early_exit(src_dev, lengths, -1.0)          # nobody leaves early: mean entropies of every exit
        thr = float(H[:3].min(dim=0).values.median())
        ee = eec.GraphedEarlyExit(m, src_dev.shape[0], src_dev.shape[2], thr)
        full = eec.GraphedForward(m, src_dev.shape[0], src_dev.shape[2])
        ee(src_dev, lengths)
        full(src_dev, lengths)

        def t(run, n=10):
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(n):
                run()
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) / n
        ms_ee, ms_full = t(ee.replay), t(full.replay)
        exit_index = ee.out[0].cpu()
    hist = [int((exit_index == e).sum()) for e in range(N_EXITS)]
    return {"model": f"splitformer {N_EXITS} exits x {layers} layers, batch {src_dev.shape[0]}", "criterion": "mean frame entropy < threshold",
            "threshold": round(thr, 4), "utterances_leaving_at_exit": hist, "ms": round(ms_ee, 3), "rtfx": round(audio_s / (ms_ee / 1e3), 1),
            "all_exits_forward_ms": round(ms_full, 3), "gpu_launches": ee.launches, "launch": "one CUDA graph replay, no host sync"}


def torch_eager_leg(layers, dev, src_dev, lengths, tg_dev, tl_dev, audio_s):
    """"the existing Blackwell kernel bar" -- the UNMODIFIED reference module (baseline/_ref: models/model/early_exit.py
    Early_conformer over torchaudio.models.Conformer, nn.CTCLoss, clip_grad_norm_, NoamOpt(AdamW); train.py:53-70) run by eager PyTorch
    on the same B200: cuBLAS / cuDNN / SDPA / ATen kernels, fp32 with torch's default TF32 settings like the reference, and once more under
    bf16 autocast.  Same batch, same weights, same step contents.  A reported baseline: none of this is on the product path."""
    ref = load_reference()
    if ref is None:
        return {"unavailable": "baseline/_ref absent (python baseline/install_ref.py in the authoring container)"}

    def t(fn, n=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / n
    res = {"what": "the reference's own Early_conformer module + train.py step (baseline/_ref, unmodified), eager PyTorch on the same GPU "
                   "(library kernels)", "kind": "reference"}
    for key, ac in (("fp32", False), ("bf16_autocast", True)):
        tr = ReferenceTrainer(ref, layers, dev, autocast=ac)
        res[f"train_ms_per_step_{key}"] = round(t(lambda: tr.step(src_dev, lengths, tg_dev, tl_dev)), 2)
        tr.model.eval()

        def infer():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                tr.model(src_dev, lengths)
        ms = t(infer)
        res[f"inference_all_exits_ms_{key}"] = round(ms, 2)
        if ac:
            res["inference_rtfx_bf16_autocast"] = round(audio_s / (ms / 1e3), 1)
        del tr
        torch.cuda.empty_cache()
    return res


def fbank_leg(dev, with_cpu):
    """waveform -> 80-dim power-mel features (util/data_loader.py:7-18) for the whole 64 x 15 s batch on the GPU
    (csrc/fbank.cu + two split-bf16 tcgen05 GEMMs), next to the numpy port of the reference's per-utterance CPU transform."""
    import eec
    g = torch.Generator().manual_seed(7)
    L = (T_IN - 1) * 160
    waves = (torch.randn(B, L, generator=g) * 0.1).to(dev)
    lens = torch.full((B,), L, dtype=torch.int64, device=dev)
    fb = eec.Fbank().cuda_tables(dev)
    for _ in range(3):
        fb(waves, lens)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        fb(waves, lens)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    leg = {"ms_per_batch": round(ms, 3), "audio_s_per_s": round(B * L / 16000.0 / (ms / 1e3), 1),
           "algorithmic_bytes": B * L * 4 + B * N_MELS * T_IN * 4, "batch": f"{B} x {L} samples (15 s at 16 kHz) -> ({B}, {N_MELS}, {T_IN})"}
    if with_cpu:
        from oracle import fbank_oracle as FO
        w1 = waves[0].cpu().numpy()
        t0 = time.perf_counter()
        FO.fbank(w1)
        leg["cpu_port_ms_per_utterance"] = round((time.perf_counter() - t0) * 1e3, 1)
    return leg


def ncu_traffic(key):
    """(bytes per launch, source) of kernel `key` from profiles/ncu_traffic.json -- written by tools/ncu_traffic.py from an
    `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum) -- or (None, None)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            e = json.load(f)[key]
        return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), e.get("source")
    except Exception:   # noqa: BLE001
        return None, None


def roofline_dominant(dev, pk):
    """Dominant kernel = the FFN up-projection GEMM (M=23936, N=2048, K=256, bias+SiLU epilogue; 24 launches per
    forward, ~37% of model FLOPs together with its twin).  Timed alone with CUDA events, L2 flushed between launches."""
    from eec import ops
    M, N, K = B * t_out(T_IN), 2048, 256
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.gemm(a, w, out, M, N, K, bias=bias, act=ops.ACT_SILU)
    total, n = 0.0, 10
    for _ in range(n):
        flush.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        ops.gemm(a, w, out, M, N, K, bias=bias, act=ops.ACT_SILU)
        ev1.record()
        torch.cuda.synchronize()
        total += ev0.elapsed_time(ev1)
    ms = total / n
    fl = 2.0 * M * N * K
    ach = fl / (ms / 1e3) / 1e12
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel, taken from the ncu --set full capture that
    # tools/ncu_traffic.py summarised into profiles/ncu_traffic.json (null when that file does not name this kernel: nothing is typed in)
    traffic, traffic_src = ncu_traffic("ffn_up_silu")
    return {"kernel": "gemm_ws2_kernel<SILU> (CTA pairs, weight-stationary) FFN up-proj 23936x2048x256 +bias+SiLU", "bound": "tensor",
            "achieved": round(ach, 2), "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_burst"], 4),
            "traffic": traffic, "traffic_source": traffic_src, "ms_per_launch": round(ms, 4), "peak_source": pk["src"] + " burst (kernel timed alone)",
            "algorithmic_flops_per_launch": fl, "algorithmic_bytes_per_launch": 2 * (M * K + N * K + M * N)}


def roofline_hbm_leg(dev, pk):
    """The north star's second target (>= 70 % of HBM peak on the norm kernels): LayerNorm backward at the BASELINE row count, timed alone
    with CUDA events and an L2 flush between launches.  Algorithmic bytes per launch: dy + x + dx in + dx out (fp32) + the bf16 operand
    copy = N x 256 x 18 B = 110 MB; 60 launches per training step (48 of them with a bf16 upstream gradient: 98 MB)."""
    from eec import ops
    N = B * t_out(T_IN)
    x = torch.randn(N, 256, device=dev)
    dy = torch.randn(N, 256, device=dev)
    g, b_ = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    out = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(N, device=dev), torch.empty(N, device=dev)
    ops.layernorm_fwd(x, g, b_, out, mean, rstd)
    dx = torch.zeros(N, 256, device=dev)
    dcopy = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
    dg, db, cs = torch.zeros(256, device=dev), torch.zeros(256, device=dev), torch.zeros(256, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    run = lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dx, True, dg, db, dcopy, cs, 1.0)   # noqa: E731
    for _ in range(3):
        run()
    total, n = 0.0, 10
    for _ in range(n):
        flush.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run()
        ev1.record()
        torch.cuda.synchronize()
        total += ev0.elapsed_time(ev1)
    ms = total / n
    nbytes = N * 256 * (4 * 4 + 2) + N * 8
    ach = nbytes / (ms / 1e3) / 1e9
    return {"kernel": "layernorm_bwd_kernel<bf16 copy> 23936 x 256 (+ dgamma/dbeta, fused bias-gradient column sums)", "bound": "hbm",
            "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(ach / pk["hbm"], 4), "traffic": ncu_traffic("layernorm_bwd")[0], "traffic_source": ncu_traffic("layernorm_bwd")[1],
            "ms_per_launch": round(ms, 4), "algorithmic_bytes_per_launch": nbytes, "launches_per_step": 60}


def port_step(sd, src, lengths, targets, tl):
    """fallback when baseline/_ref is absent: the oracle port of the same step (kind = "port"; forward + loss + backward only)"""
    from oracle import conformer_oracle as O
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and k != "positional_encoder.pe" else v)
           for k, v in sd.items()}
    out = O.early_conformer_forward(sdg, src, lengths, training=True, bn_out={})
    Tn = out.shape[2]
    in_len = torch.full((out.shape[1],), Tn, dtype=torch.long)
    loss = sum(torch.nn.functional.ctc_loss(out[e].permute(1, 0, 2), targets, in_len, tl, blank=0, zero_infinity=True)
               for e in range(out.shape[0]))
    loss.backward()
    return float(loss.detach())


def cpu_arm(layers, sample_b):
    """-> (step(), kind, infer()): the reference's CPU path on `sample_b` utterances of the benchmark batch -- the real reference
    module from baseline/_ref when present (kind "reference"), else the oracle port (kind "port")."""
    src, lengths, targets, tl = synthetic_batch(sample_b, 1234)
    ref = load_reference()
    if ref is not None:
        tr = ReferenceTrainer(ref, layers, torch.device("cpu"))

        def infer():
            tr.model.eval()
            with torch.no_grad():
                tr.model(src, lengths)
            tr.model.train()
        return (lambda: tr.step(src, lengths, targets, tl)), "reference", infer, lengths
    sd = synthetic_state_dict(layers)
    from oracle import conformer_oracle as O

    def infer():
        with torch.no_grad():
            O.early_conformer_forward(sd, src, lengths)
    return (lambda: port_step(sd, src, lengths, targets, tl)), "port", infer, lengths


def cpu_baseline(layers, sample_b=16, steps=2):
    """The reference's CPU path on the GPU box's host cores, on a bounded sample (`sample_b` of the 64 utterances per step: 16 is the
    sub-batch size the reference's own loader feeds the model, batch_size 64 / n_batch_split 4, util/data_loader.py:166-188)."""
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step, kind, infer, lengths = cpu_arm(layers, sample_b)
    step()                                        # one untimed step (allocator, thread pool)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    # BASELINE configs[0] (the reference's CPU-runnable case): eval forward of all six exits, no grad, same sample
    t1 = time.perf_counter()
    infer()
    dti = time.perf_counter() - t1
    audio_s = float(lengths.sum()) * FRAME_S
    return {"value": round(sample_b / dt, 3), "unit": "utt/s", "cores": cores, "kind": kind,
            "sample": f"{steps} training steps (fwd + 6-exit CTC + bwd + clip + Noam/AdamW, fp32) on {sample_b} of the 64 utterances per "
                      f"step, same T_in={T_IN}, {N_EXITS}x{layers} layers; {dt:.2f} s per step",
            "inference_rtfx_all_exits": round(audio_s / dti, 1),
            "inference_sample": f"eval forward of all {N_EXITS} exits (fp32, no grad) on the same {sample_b} utterances: {dti:.2f} s for {audio_s:.0f} s of audio"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step (baseline/_ref, unmodified) on all host cores, each step a
    bounded sample (--cpu-sample utterances of the 64-utterance batch) of the same workload, same metric / unit / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    layers = args.layers_per_exit
    sb = args.cpu_sample
    step, kind, _, _ = cpu_arm(layers, sb)
    warm = args.warmup
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sb / dt
    line = {
        "impl": "reference", "metric": "train_utts_per_sec", "value": round(value, 3), "unit": "utt/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(layers),
        "cpu_baseline": {"value": round(value, 3), "unit": "utt/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} timed steps x {sb} utterances (of the 64-utterance batch; 16 = the reference loader's own "
                                   f"sub-batch, batch_size 64 / n_batch_split 4), {dt:.2f} s per step"},
        "e2e": {"value": round(value, 3), "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


_JSON_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--layers-per-exit", type=int, default=2, help="2 = BASELINE configs[1]; 3 = configs[2] (18 layers)")
    ap.add_argument("--cpu-sample", type=int, default=16, help="utterances per CPU step of the reference arm / cpu_baseline leg (16 = the reference loader's own sub-batch; 64 = the full batch)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-opt", action="store_true", help="time forward + loss + backward only (no clip / Noam / AdamW update)")
    ap.add_argument("--no-graph", action="store_true", help="issue the ~600 kernels of a step eagerly instead of replaying the CUDA graph")
    ap.add_argument("--skip-rtfx", action="store_true")
    ap.add_argument("--train-drop-prob", type=float, default=0.0, help="dropout probability of the HEADLINE step (default 0, the parity "
                    "configuration of SURVEY 8d); used with --profile to capture the launch list of the dropout step")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one all-reduce of the flat gradient buffer after backward "
                    "(outside the graph) instead of per-exit-group all-reduces overlapped with backward")
    ap.add_argument("--drop-prob", type=float, default=0.1, help="extra leg: the same training step with dropout at this probability "
                    "(the reference's default); the headline step runs at 0 like the parity tests (SURVEY 8d). 0 skips the leg")
    ap.add_argument("--skip-parity", action="store_true", help="skip the in-run parity check of the headline step against the reference's CPU forward")
    ap.add_argument("--skip-deep", action="store_true", help="skip the 18-layer (BASELINE configs[2]) leg")
    ap.add_argument("--skip-aed", action="store_true", help="skip the AED-mode (BASELINE configs[4]) leg")
    ap.add_argument("--grad-dtype", default="fp32", choices=["fp32", "bf16"], help="N > 1: wire format of the overlapped gradient all-reduce")
    ap.add_argument("--kernel-trace", default="", help="write a torch.profiler kernel list of one step (rank 0) to this path")
    ap.add_argument("--profile", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (for ncu "
                    "--profile-from-start off) and skip the e2e / rtfx / roofline / cpu legs")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner, ...) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
