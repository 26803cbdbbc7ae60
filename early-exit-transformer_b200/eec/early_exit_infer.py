"""Dynamic early-exit inference with on-device batch compaction (north star).
NOT in the reference (its inference.py evaluates every exit, :65-82): the oracle for this path is
oracle.conformer_oracle.early_exit_select on the reference's full output ("parity unpinned")."""
from __future__ import annotations

import torch

from . import engine, ops

D, V = engine.D, engine.V


def run(model, src, lengths, threshold: float):
    cfg = model._cfg()
    P = model._tensor_dict()
    W = model._operands
    src = src.contiguous().float()
    dev, f32, i32 = src.device, torch.float32, torch.int32
    B = src.shape[0]
    x, T = engine.frontend_forward(P, W, src, cfg, None)
    engine.check_lengths(lengths, T)
    N = B * T
    lengths_dev = lengths.to(device=dev, dtype=torch.int64, non_blocking=True)
    key_len = torch.empty(B, dtype=i32, device=dev)
    ops.encoder_lengths(lengths_dev, key_len, T, 4, 0)
    E = cfg.n_exits
    # device-side bookkeeping (ping-pong)
    row_map = torch.arange(B, dtype=i32, device=dev)
    row_map2 = torch.empty_like(row_map)
    key_len2 = torch.empty_like(key_len)
    gather_idx = torch.empty(B, dtype=i32, device=dev)
    n_alive = torch.full((1,), B, dtype=i32, device=dev)
    exit_index = torch.full((B,), -1, dtype=i32, device=dev)
    tokens = torch.full((B, T), -1, dtype=i32, device=dev)
    n_tokens = torch.zeros(B, dtype=i32, device=dev)
    mean_ent = torch.full((E, B), float("nan"), dtype=f32, device=dev)
    lp = torch.empty(N, V, dtype=f32, device=dev)
    am = torch.empty(N, dtype=i32, device=dev)
    en = torch.empty(N, dtype=f32, device=dev)
    logits_ws = torch.empty(N, V, dtype=f32, device=dev) if cfg.precision == "fp32" else None
    # the compaction target starts as zeros: rows behind the survivors must always hold finite numbers (see PAD below)
    x2 = torch.zeros_like(x)
    # one padding utterance behind the survivors is still computed (on stale but finite data): 128-row K/V tiles of the
    # tensor-core attention overhang into it; more when an utterance is shorter than a tile
    PAD = max(1, -(-128 // max(T // 2, 1)))
    T2, pad = (T + 1) // 2, T % 2
    len_alive = lengths_dev                                   # raw fbank lengths in compacted order (Splitformer branch mask)
    try:
        for e in range(E):
            # from the second group on only the survivors are computed: the kernels read n_alive on the device
            # (eec_set_active_items), so the finished utterances' rows cost nothing and nothing syncs with the host
            if e > 0:
                ops.set_active_items(n_alive, T, PAD)
            x_in = x
            for l in range(cfg.n_layers):
                x = engine.layer_forward(P, W, f"conformer.{e}.conformer_layers.{l}.", x, key_len, B, T, cfg, False, None)
            if cfg.splitformer and (e == 0 or e == E - 1):
                # early_exit.py:314-356 on the surviving rows: stride-2 branch on the group's input, raw-length key mask
                i = e // (E - 1)
                xd = torch.empty(B * T2, D, dtype=f32, device=dev)
                ops.stride2_gather(x_in, xd, B, T)
                if e > 0:
                    len_alive = torch.empty_like(lengths_dev)
                    ops.gather_i64(lengths_dev, row_map, len_alive)
                    ops.set_active_items(n_alive, T2, PAD)
                len2 = torch.empty(B, dtype=i32, device=dev)
                ops.encoder_lengths(len_alive, len2, T2, 2, pad)
                yd = engine.layer_forward(P, W, f"conformer_parallel.{i}.conformer_layers.0.", xd, len2, B, T2, cfg, False, None)
                if x is x_in:
                    x = x.clone()
                ops.repeat2_add(yd, x, B, T)
                if e > 0:
                    ops.set_active_items(n_alive, T, PAD)
            xh = engine.to_act(x, cfg)
            Wh = W.get(f"linears.{e}.weight", P[f"linears.{e}.weight"], (V, D))
            ops.head_logsoftmax(xh, Wh, P[f"linears.{e}.bias"], lp, am, en, logits_ws)
            ops.exit_select(en, am, key_len, row_map, n_alive, e, e == E - 1, float(threshold), exit_index, tokens, n_tokens,
                            row_map2, key_len2, gather_idx, mean_ent, B, T)
            if e < E - 1:
                ops.gather_rows(x, x2, gather_idx, n_alive, B, T * D)
                x, x2 = x2, x
                row_map, row_map2 = row_map2, row_map
                key_len, key_len2 = key_len2, key_len
    finally:
        ops.set_active_items(None)
    return exit_index, tokens, n_tokens, mean_ent
