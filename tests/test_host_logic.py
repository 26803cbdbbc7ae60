"""CPU tests: C-ABI surface (load + every declared symbol, no compute), host-side mirror of the reference
interface, and the world_size-2 data-parallel exchange step over gloo."""
import ctypes
import os
import re
import socket

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import eec
    lib = eec.load()
    header = open(os.path.join(ROOT, "include", "eec.h")).read()
    declared = set(re.findall(r"\b(eec_[a-z0-9_]+)\s*\(", header))
    declared -= {"eec_gemm_desc"}
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/eec.h but not exported by libeec.so"
    assert set(eec.EXPORTS) <= declared
    assert lib.eec_version() >= 100
    assert isinstance(lib.eec_last_error(), bytes)
    # struct layout the Python side marshals must match the C side (size check against the header's field list)
    from eec.lib import GemmDesc
    assert ctypes.sizeof(GemmDesc) % 8 == 0 and ctypes.sizeof(GemmDesc) >= 200


def test_module_mirror_matches_reference_state_dict_layout():
    import eec
    from oracle import conformer_oracle as O
    kw = dict(src_pad_idx=0, n_enc_exits=6, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
              d_feed_forward=2048, n_enc_layers=2, features_length=80, drop_prob=0.1, depthwise_kernel_size=31, device="cpu")
    m = eec.Early_conformer(**kw)
    sd = m.state_dict()
    assert len(sd) == 413 and sum(p.numel() for p in m.parameters()) == 31536128
    ref = O.make_params(0)
    assert set(ref) == set(sd) and all(tuple(ref[k].shape) == tuple(sd[k].shape) for k in sd)
    assert sd["conformer.0.conformer_layers.0.conv_module.sequential.3.num_batches_tracked"].dtype == torch.int64
    s = eec.Splitformer(**kw)
    assert len(s.state_dict()) == 479
    with pytest.raises(eec.EecError):
        m.eval()(torch.zeros(1, 80, 100), torch.tensor([100]))    # no CPU fallback: fails loudly
    with pytest.raises(ValueError):
        eec.Early_conformer(**{**kw, "depthwise_kernel_size": 30})
    with pytest.raises(AssertionError):
        eec.Early_conformer(**{**kw, "n_head": 6})


def test_same_seed_same_init_as_torch_containers():
    """Construction order mirrors the reference, so a fixed seed gives identical default weights."""
    import eec
    kw = dict(src_pad_idx=0, n_enc_exits=2, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=100,
              d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.0, depthwise_kernel_size=31, device="cpu")
    torch.manual_seed(0)
    a = eec.Early_conformer(**kw).state_dict()
    torch.manual_seed(0)
    b = eec.Early_conformer(**kw).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    torchaudio = pytest.importorskip("torchaudio")
    torch.manual_seed(0)
    # the reference's constructor sequence: conv_subsample, pos-enc, linears, then torchaudio Conformers
    _ = [torch.nn.Conv1d(80, 256, 3, 2), torch.nn.Conv1d(256, 256, 3, 2)]
    _ = [torch.nn.Linear(256, 256) for _ in range(2)]
    ta = torchaudio.models.Conformer(input_dim=256, num_heads=8, ffn_dim=2048, num_layers=1, depthwise_conv_kernel_size=31)
    tsd = ta.state_dict()
    for k, v in tsd.items():
        assert torch.equal(a["conformer.0." + k], v), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
    import torch.distributed as dist
    from eec import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a model whose grads live in one flat buffer, exactly as engine.model_backward leaves them
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.LayerNorm(4))
    if rank == 1:
        for p in model.parameters():
            p.data.add_(1.0)
    D.broadcast_parameters(model, 0)
    names = [n for n, _ in model.named_parameters()]
    total = sum(p.numel() for p in model.parameters())
    flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
    off = 0
    for p in model.parameters():
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    model.__dict__["_flat_grad"] = flat
    D.all_reduce_gradients(model)
    lo, hi = D.shard_range(7, rank, world)
    q.put((rank, flat.tolist(), [p.detach().reshape(-1).tolist() for p in model.parameters()], (lo, hi), names))
    dist.destroy_process_group()


def test_data_parallel_exchange_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    flat0, flat1 = torch.tensor(res[0][1]), torch.tensor(res[1][1])
    expect = torch.arange(flat0.numel(), dtype=torch.float32) * 1.5      # mean of 1x and 2x
    assert torch.allclose(flat0, expect) and torch.allclose(flat1, expect)
    for a, b in zip(res[0][2], res[1][2]):
        assert a == b                                                     # broadcast made the replicas identical
    assert res[0][3] == (0, 4) and res[1][3] == (4, 7)                    # utterance sharding covers [0,7) once


def test_shard_batch_partitions_every_utterance_once():
    from eec import distributed as D
    x = torch.arange(10)
    parts = [D.shard_batch([x], r, 4)[0] for r in range(4)]
    assert torch.equal(torch.cat(parts), x) and max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_full_conformer_mirror_layout_and_init():
    """eec.full_conformer (AED mode) instantiates the reference's modules in the reference's order:
    123-entry state_dict for 2 exits x 1 layer x 1 decoder layer, decoder / embedding parameters reproducible from the
    seed (the golden fixture relies on this), constructor errors preserved."""
    import eec
    kw = dict(trg_pad_idx=126, n_enc_exits=2, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
              d_feed_forward=2048, n_enc_layers=1, n_dec_layers=1, features_length=80, drop_prob=0.0, depthwise_kernel_size=31,
              device=torch.device("cpu"))
    torch.manual_seed(5)
    a = eec.full_conformer(**kw)
    torch.manual_seed(5)
    b = eec.full_conformer(**kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert len(sa) == 123 and list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    keys = list(sa)
    assert keys[0] == "layer_norm.weight" and keys[2] == "emb.weight" and "linears_1.0.weight" in sa and "linears_2.1.bias" in sa
    assert "positional_encoder_1.pe" in sa and "decoders.1.layers.0.multihead_attn.in_proj_weight" in sa
    assert a._param_names[0] == "conv_subsample.sequential.0.weight" and "linears.0.weight" in a._param_names
    assert not any(n.startswith(("decoders", "emb", "linears_2")) for n in a._param_names)
    with pytest.raises(AssertionError):
        eec.full_conformer(**{**kw, "n_head": 6})
    with pytest.raises(Exception):   # no CPU path: the encoder half refuses CPU tensors
        a(torch.zeros(1, 80, 163), torch.tensor([163]), torch.zeros(1, 4, dtype=torch.long))


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="the reference tree exists in the authoring container only")
def test_dropin_rebinds_reference_classes():
    """eec.dropin.install(): the unmodified reference package resolves Early_conformer / Splitformer /
    full_conformer to the eec classes, which accept the reference's keyword call (train.py:166-178) and keep its layout."""
    import importlib
    import sys
    import eec
    import eec.dropin
    sys.path.insert(0, "/root/reference")
    try:
        ref = eec.dropin.install()
        assert ref.Early_conformer is eec.Early_conformer and ref.Splitformer is eec.Splitformer and ref.full_conformer is eec.full_conformer
        m = ref.Early_conformer(src_pad_idx=0, n_enc_exits=2, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
                                d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.1, depthwise_kernel_size=31,
                                device=torch.device("cpu"))
        assert len(m.state_dict()) == 1 + 4 + 2 * 2 + 2 * 33   # pe + front end + heads + 2 layers x 33 entries (413 at 12L/6E)
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        importlib.invalidate_caches()


def test_dropin_ctc_patch_is_scoped_to_one_module():
    """eec.dropin.patch_ctc(train_module): the module's `nn.CTCLoss(blank=0, zero_infinity=True)` builds eec.CTCLoss; torch.nn itself and
    every other configuration are untouched."""
    import types
    import eec
    import eec.dropin
    mod = types.ModuleType("fake_train")
    mod.nn = torch.nn
    orig = torch.nn.CTCLoss
    eec.dropin.patch_ctc(mod)
    assert torch.nn.CTCLoss is orig                                           # global namespace unchanged
    assert isinstance(mod.nn.CTCLoss(blank=0, zero_infinity=True), eec.CTCLoss)
    assert isinstance(mod.nn.CTCLoss(blank=0), orig) and isinstance(torch.nn.CTCLoss(blank=0, zero_infinity=True), orig)
    assert mod.nn.Linear is torch.nn.Linear and mod.nn.functional is torch.nn.functional
    with pytest.raises(AttributeError):
        eec.dropin.patch_ctc(types.ModuleType("no_nn"))


def _overlap_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "early-exit-transformer_b200"))
    import torch.distributed as dist
    import eec
    from eec import distributed as D, engine
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(src_pad_idx=0, n_enc_exits=3, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
              d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.0, depthwise_kernel_size=31, device="cpu")
    m = eec.Splitformer(**kw)       # host mirror only (no kernels run on the CPU): parameter layout of the flat gradient buffer
    names = m._param_names
    P = m._tensor_dict()
    spans, total = engine.group_ranges(P, names, 3)
    flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
    red = D.OverlappedGradReducer(m)
    seen = []
    # replay the order engine.model_backward issues: groups last to first, then what is not an exit group
    for e in reversed(range(3)):
        red.on_ready(flat, *spans[e])
        seen.append(spans[e])
    lo0, hi0 = spans[0][0], spans[-1][1]
    red.on_ready(flat, 0, lo0)
    red.on_ready(flat, hi0, total)
    red.finish()
    D.all_reduce_gradients(m)       # must be a no-op now (would average a second time otherwise: same values, so check calls)
    q.put((rank, total, spans, (lo0, hi0), red.calls, bool(torch.allclose(flat, torch.arange(total, dtype=torch.float32) * 1.5))))
    dist.destroy_process_group()


def test_overlapped_grad_reducer_world2_gloo():
    """Per-exit-group all-reduce (eec.distributed.OverlappedGradReducer) covers the flat gradient buffer exactly once."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, total, spans, (lo0, hi0), calls, ok in res:
        assert ok and calls == 5
        assert spans[0][1] == spans[1][0] and spans[1][1] == spans[2][0]          # groups are contiguous and ordered
        assert lo0 > 0 and hi0 < total                                               # front end + heads before, Splitformer branches after
        assert all(hi - lo == spans[0][1] - spans[0][0] for lo, hi in spans)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_utts_per_sec" and d["unit"] == "utt/s" and d["higher_is_better"] is True
    # the real reference module (baseline/_ref, installed by baseline/install_ref.py) when present, else the oracle port
    kind = "reference" if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "models")) else "port"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_length_sorted_split_matches_reference_collate():
    """eec.batching.length_sorted_split == the chunking loop of CollatePaddingFn.__call__ (util/data_loader.py:163-188), restated here
    statement for statement on (length, id) records, for many random mini-batches; plus the properties train.py:22-26 relies on."""
    import random
    from eec import batching

    def reference_chunks(batch, n_split):          # batch: [(width, uid)]; x[0].size(1) -> x[0]
        batch = sorted(batch, key=lambda x: x[0], reverse=True)
        s_sum = sum(x[0] for x in batch) / n_split
        p_sum, chunked_batch, init, end, p_split = 0, [], 0, 0, 0
        for w, *_ in batch:
            p_sum += w
            if p_sum >= s_sum:
                chunked_batch.append(batch[init:end + 1])
                p_sum = 0
                p_split += 1
                init = end + 1
            end += 1
        if p_split != n_split:
            chunked_batch.append(batch[init:end])
        return [[u for _, u in c] for c in chunked_batch]

    rng = random.Random(5)
    n_ok = 0
    for trial in range(300):
        n = rng.randint(1, 70)
        n_split = rng.choice([1, 2, 4, 8])
        lengths = [rng.randint(200, 3200) for _ in range(n)]
        if trial % 7 == 0:
            lengths = [lengths[0]] * n                               # all equal: ties keep the input order
        got = batching.length_sorted_split(lengths, n_split)
        assert got == reference_chunks([(w, i) for i, w in enumerate(lengths)], n_split)
        flat = [i for c in got for i in c]
        assert sorted(flat) == list(range(n))                        # every utterance exactly once
        assert all(lengths[a] >= lengths[b] for a, b in zip(flat, flat[1:]))
        n_ok += batching.trains_on(got, n_split)
    assert n_ok > 200
    # the BASELINE shape: 64 utterances, 4 chunks -> ~16 per chunk, far less padding than one 64-utterance call
    lengths = [rng.randint(750, 1501) for _ in range(64)]
    ch = batching.length_sorted_split(lengths, 4)
    assert len(ch) == 4 and 10 <= min(map(len, ch)) and max(map(len, ch)) <= 24
    assert batching.padding_ratio(lengths, ch) < 0.5 * batching.padding_ratio(lengths, [list(range(64))])
    # graph cache policy: exact (B, T_in), bucketed target width, LRU eviction
    made = []
    cache = batching.GraphedStepCache(lambda B, T, L: made.append((B, T, L)) or (B, T, L), capacity=2)
    assert cache.get(16, 1501, 70) == (16, 1501, 80) and cache.get(16, 1501, 75) == (16, 1501, 80) and cache.hits == 1
    cache.get(17, 1501, 70); cache.get(16, 1400, 70)
    assert cache.evictions == 1 and cache.captures == 3 and (16, 1501, 80) not in cache.steps
