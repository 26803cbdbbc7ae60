"""GPU parity of the whole hot path (through eec.Early_conformer / Splitformer -> C ABI -> sm_100a kernels)
against (a) the committed golden vectors produced by the real reference and (b) the CPU oracle run live.
Tolerances are the north star's: 1e-3 relative (maxabs(a-b)/maxabs(b)) for fp32, 2e-2 for bf16, greedy CTC
tokens bit-exact for fp32."""
import os

import numpy as np
import pytest
import torch

from oracle import conformer_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["ec_e2l1_b3_t163", "ec_e3l2_b4_t331", "sf_e2l1_b3_t166", "sf_e3l1_b2_t203"]
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build(name, g, precision):
    import eec
    split = name.startswith("sf")
    cls = eec.Splitformer if split else eec.Early_conformer
    m = cls(src_pad_idx=0, n_enc_exits=int(g["n_exits"]), enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
            max_len=2000, d_feed_forward=2048, n_enc_layers=int(g["n_layers"]), features_length=80, drop_prob=0.0,
            depthwise_kernel_size=31, device=torch.device("cuda"))
    seed = int(g["seed"])
    sd = O.make_params(seed, n_exits=int(g["n_exits"]), n_layers=int(g["n_layers"]), splitformer=split)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda")
    m.precision = precision
    src, lengths = O.synthetic_batch(int(g["B"]), int(g["t_in"]), seed=seed + 1)
    return m, sd, src, lengths


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_eval_logprobs_vs_reference_golden(name, precision):
    import eec
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, precision)
    m.eval()
    with torch.no_grad():
        out = m(src.cuda(), lengths)
    ref = torch.from_numpy(g["eval_logprobs"])
    assert out.shape == ref.shape and out.dtype == torch.float32
    for e in range(ref.shape[0]):
        assert rel(out[e], ref[e]) < TOL[precision], (e, rel(out[e], ref[e]))
    if precision == "fp32":
        tokens, n_tok = eec.greedy_decode(out)
        tokens, n_tok = tokens.cpu(), n_tok.cpu()
        assert np.array_equal(n_tok.numpy(), g["greedy_counts"])  # bit-exact greedy CTC (north star)
        flat = [int(t) for e in range(ref.shape[0]) for b in range(ref.shape[1]) for t in tokens[e, b, : int(n_tok[e, b])]]
        assert np.array_equal(np.array(flat, dtype=np.int64), g["greedy_flat"])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_train_step_vs_reference_golden(name, precision):
    import eec
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, precision)
    m.train()
    out = m(src.cuda(), lengths)
    ref = torch.from_numpy(g["train_logprobs"])
    for e in range(ref.shape[0]):
        assert rel(out[e], ref[e]) < TOL[precision]
    targets, tl = torch.from_numpy(g["targets"]), torch.from_numpy(g["target_lengths"])
    # (a) the reference's own call shape (train.py:57-63)
    ctc = eec.CTCLoss(blank=0, zero_infinity=True)
    in_len = torch.full((out.size(1),), out.size(2), dtype=torch.long)
    loss = sum(ctc(enc.permute(1, 0, 2), targets, in_len, tl) for enc in out)
    lt = TOL[precision]
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < lt
    m.zero_grad()
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    P = dict(m.named_parameters())
    gt = 5e-3 if precision == "fp32" else 6e-2
    worst = 0.0
    for n, refn in zip(names, g["grad_norms"]):
        got = P[n].grad.double().norm().item()
        err = abs(got - refn) / max(refn, 1e-3 * float(np.max(g["grad_norms"])))
        worst = max(worst, err)
        assert err < gt, (n, got, refn)
    for k in g.files:
        if k.startswith("grad::"):
            r = torch.from_numpy(g[k])
            if float(r.abs().max()) < 1e-4:  # e.g. dw-conv bias under train BN: pure round-off
                continue
            assert rel(P[k[6:]].grad, r) < gt, k
    if precision == "fp32":
        msd = m.state_dict()
        bn = "conformer.0.conformer_layers.0.conv_module.sequential.3."
        assert np.allclose(msd[bn + "running_mean"].cpu().numpy(), g["bn_running_mean"], rtol=1e-3, atol=1e-5)
        assert np.allclose(msd[bn + "running_var"].cpu().numpy(), g["bn_running_var"], rtol=1e-3, atol=1e-5)
        assert int(msd[bn + "num_batches_tracked"]) == int(g["bn_num_batches_tracked"])
    # (b) fused multi-exit CTC: same value in one launch
    m2, _, _, _ = build(name, g, precision)
    m2.train()
    out2 = m2(src.cuda(), lengths)
    loss2 = eec.multi_exit_ctc_loss(out2, targets, tl)
    assert abs(loss2.item() - loss.item()) < 1e-4 * abs(loss.item())
    loss2.backward()
    P2 = dict(m2.named_parameters())
    k = "conformer.0.conformer_layers.0.ffn1.sequential.1.weight"
    # (the fused loss hands the encoder d(loss)/d(logits) directly, the per-exit calls go through the general log-softmax backward: equal up to
    #  fp32 round-off, which the bf16 operand casts amplify to a few bf16 ulps)
    assert rel(P2[k].grad, P[k].grad) < (1e-3 if precision == "fp32" else 5e-3)


def test_ctc_known_answers():
    import eec
    g = np.load(os.path.join(GOLDEN, "ctc_kat.npz"))
    lp = torch.from_numpy(g["lp"]).cuda().requires_grad_(True)  # (B,T,V)
    loss = eec.multi_exit_ctc_loss(lp.unsqueeze(0), torch.from_numpy(g["targets"]), torch.from_numpy(g["target_lengths"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    ref = torch.from_numpy(g["grad"])
    assert float((lp.grad.cpu() - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    assert float(lp.grad[2].abs().max()) == 0.0  # infeasible utterance: zero_infinity


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ragged_batch_vs_live_oracle(precision):
    """Bigger ragged batch, 2 exits x 2 layers, checked against the CPU oracle run in this process."""
    import eec
    sd = O.make_params(7, n_exits=2, n_layers=2)
    src, lengths = O.synthetic_batch(6, 419, seed=8, min_frac=0.3)
    lengths[3] = 3  # length//4 == 0 -> every key masked for this utterance
    m = eec.Early_conformer(src_pad_idx=0, n_enc_exits=2, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
                            max_len=2000, d_feed_forward=2048, n_enc_layers=2, features_length=80, drop_prob=0.1,
                            depthwise_kernel_size=31, device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = precision
    with torch.no_grad():
        out = m(src.cuda(), lengths)
        ref = O.early_conformer_forward(sd, src, lengths)
    for e in range(2):
        assert rel(out[e], ref[e]) < TOL[precision]
    assert torch.isfinite(out).all()
    # the reference's precondition
    with pytest.raises(AssertionError):
        m(src.cuda(), torch.full_like(lengths, 100))


def test_early_exit_matches_oracle_selection():
    import eec
    sd = O.make_params(11, n_exits=3, n_layers=1)
    src, lengths = O.synthetic_batch(5, 331, seed=12, min_frac=0.4)
    m = eec.Early_conformer(src_pad_idx=0, n_enc_exits=3, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
                            max_len=2000, d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.0,
                            depthwise_kernel_size=31, device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        full = m(src.cuda(), lengths).cpu()
    T = full.shape[2]
    key_len = O.encoder_lengths(lengths, T)
    _, _, Hm = O.early_exit_select(full, key_len, 1e9)
    thr = float(Hm[0].median())  # some utterances leave at exit 0, others continue
    ex_ref, tok_ref, _ = O.early_exit_select(full, key_len, thr)
    exit_index, tokens, n_tokens, mean_ent = m.forward_early_exit(src.cuda(), lengths, thr)
    assert exit_index.cpu().tolist() == ex_ref.tolist()
    assert len(set(ex_ref.tolist())) > 1
    for b in range(5):
        assert tokens[b, : int(n_tokens[b])].cpu().tolist() == tok_ref[b]
    # survivors' outputs are unchanged by compaction: entropies agree with the full-batch run
    for b in range(5):
        for e in range(int(ex_ref[b]) + 1):
            assert abs(float(mean_ent[e, b]) - float(Hm[e, b])) < 1e-3


def test_state_dict_roundtrip_and_errors():
    import eec
    kw = dict(src_pad_idx=0, n_enc_exits=6, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
              d_feed_forward=2048, n_enc_layers=2, features_length=80, drop_prob=0.1, depthwise_kernel_size=31, device="cuda")
    m = eec.Early_conformer(**kw)
    assert len(m.state_dict()) == 413
    assert sum(p.numel() for p in m.parameters()) == 31536128
    with pytest.raises(ValueError):
        eec.Early_conformer(**{**kw, "depthwise_kernel_size": 30})
    with pytest.raises(AssertionError):
        eec.Early_conformer(**{**kw, "n_head": 6})
    with pytest.raises(eec.EecError):
        m.eval()(torch.zeros(1, 80, 100), torch.tensor([100]))  # CPU input: no fallback
    out = m.cuda().train()(torch.zeros(1, 80, 100, device="cuda"), torch.tensor([100]))  # dropout > 0 in training runs (tests/test_gpu_dropout.py)
    assert out.shape == (6, 1, 24, 256) and torch.isfinite(out).all()
    with pytest.raises(ValueError):
        eec.Early_conformer(**{**kw, "drop_prob": 1.5}).cuda().train()(torch.zeros(1, 80, 100, device="cuda"), torch.tensor([100]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_train_step_matches_eager(precision):
    """The CUDA-graph replay of a training step (eec.GraphedTrainStep) gives the eager step's loss and gradients,
    sees new inputs and updated weights between replays, and raises the reference's length precondition on the host."""
    import eec
    name = "ec_e2l1_b3_t163"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, precision)
    m.train()
    Bn = src.shape[0]
    targets, tl = O.synthetic_targets(Bn, seed=11, lo=3, hi=6)

    def eager(x, ln, tg, tlen):
        m.zero_grad(set_to_none=True)
        out = m(x.cuda(), ln)
        loss = eec.multi_exit_ctc_loss(out, tg, tlen)
        loss.backward()
        return float(loss.detach()), {n: p.grad.detach().clone() for n, p in m.named_parameters()}

    ref_loss, ref_g = eager(src, lengths, targets, tl)
    step = eec.GraphedTrainStep(m, Bn, src.shape[2], targets.shape[1] + 3)   # wider target buffer than this batch: padded
    assert step.launches_per_step > 50
    loss = step(src, lengths, targets, tl)
    torch.cuda.synchronize()
    tol = 1e-5 if precision == "fp32" else 1e-2   # same kernels, same order: only atomics reorder
    assert abs(float(loss) - ref_loss) <= tol * abs(ref_loss)
    gmax = max(float(v.abs().max()) for v in ref_g.values())
    for n, p in m.named_parameters():   # (gradients that are analytically zero, e.g. the key bias, are compared on the global scale)
        err = float((p.grad - ref_g[n]).abs().max())
        assert err <= max(tol, 2e-4) * max(float(ref_g[n].abs().max()), 1e-3 * gmax), (n, err)
    assert m._flat_grad.data_ptr() == next(m.parameters()).grad.data_ptr()
    # new batch + changed weights: the replay must see both
    src2, lengths2 = O.synthetic_batch(Bn, src.shape[2], seed=77)
    targets2, tl2 = O.synthetic_targets(Bn, seed=12, lo=3, hi=6)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
    ref_loss2, ref_g2 = eager(src2, lengths2, targets2, tl2)
    loss2 = step(src2, lengths2, targets2, tl2)
    torch.cuda.synchronize()
    assert abs(float(loss2) - ref_loss2) <= tol * abs(ref_loss2)
    k = "conformer.0.conformer_layers.0.ffn1.sequential.1.weight"
    assert rel(dict(m.named_parameters())[k].grad, ref_g2[k]) < max(tol, 2e-4)
    # prefetch path: features uploaded on a side stream into a staging buffer, committed before the replay
    step.load_inputs(src, lengths2, targets2, tl2)           # (static src now holds the OTHER batch)
    step.prefetch(src2.pin_memory())
    step.commit_prefetch()
    loss3 = step.replay()
    torch.cuda.synchronize()
    assert abs(float(loss3) - ref_loss2) <= tol * abs(ref_loss2)
    with pytest.raises(AssertionError):
        step(src2, torch.full((Bn,), 100, dtype=torch.int64), targets2, tl2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("fixture", ["fc_e2l1d1_b2_t163", "fc_e2l1d2_b3_t203"])
def test_full_conformer_aed_vs_reference_golden(fixture, precision):
    """SURVEY §8 rows a17 + N4 / BASELINE configs[4]: eec.full_conformer (encoder AND attention-decoder stacks on the sm_100a kernels)
    against the real reference's outputs: CTC log-probs, decoder logits, `_encoder_` / `_decoder_`, the AED loss (0.7 CE + 0.3 CTC,
    train.py:44-51) and the gradient norm of EVERY parameter (encoder, decoder stacks, embedding, shared final LayerNorm).  The second
    fixture has two decoder layers per stack and ragged targets (padding inside `trg`: the key-padding mask matters)."""
    import eec
    g = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    seed, B = int(g["seed"]), int(g["B"])
    kw = dict(trg_pad_idx=126, n_enc_exits=int(g["n_exits"]), enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
              max_len=2000, d_feed_forward=2048, n_enc_layers=int(g["n_layers"]), n_dec_layers=int(g["n_dec"]), features_length=80,
              drop_prob=0.0, depthwise_kernel_size=31, device=torch.device("cuda"))
    torch.manual_seed(seed)                      # decoder / embedding parameters: the reference's default init under this seed
    m = eec.full_conformer(**kw)
    enc_sd = {}
    for k, v in O.make_params(seed, n_exits=kw["n_enc_exits"], n_layers=kw["n_enc_layers"]).items():
        k = k.replace("linears.", "linears_1.", 1) if k.startswith("linears.") else k
        k = k.replace("positional_encoder.", "positional_encoder_1.", 1) if k.startswith("positional_encoder.") else k
        enc_sd[k] = v
    r = m.load_state_dict(enc_sd, strict=False)
    assert not r.unexpected_keys
    m = m.to("cuda")
    m.precision = precision
    src, lengths = O.synthetic_batch(B, int(g["t_in"]), seed=seed + 1)
    targets, tl = torch.from_numpy(g["targets"]), torch.from_numpy(g["target_lengths"])
    trg, trg_expect = targets[:, :-1].cuda(), targets[:, 1:].cuda()
    tol = TOL[precision]
    m.eval()
    with torch.no_grad():
        dec_out, enc_out = m(src.cuda(), lengths, trg)
        enc1 = m._encoder_(src.cuda(), lengths, 1)
        dec1 = m._decoder_(trg, enc1, 1)
    assert rel(enc_out, torch.from_numpy(g["eval_enc_out"])) < tol
    assert rel(dec_out, torch.from_numpy(g["eval_dec_out"])) < tol
    assert rel(enc1, torch.from_numpy(g["encoder_1"])) < tol
    assert rel(dec1, torch.from_numpy(g["decoder_1"])) < tol
    m.train()
    att_dec, encoder = m(src.cuda(), lengths, trg)
    ctc = eec.CTCLoss(blank=0, zero_infinity=True)
    ce = torch.nn.CrossEntropyLoss()
    in_len = torch.full((B,), encoder.size(2), dtype=torch.long)
    loss_ctc = sum(ctc(enc.permute(1, 0, 2), targets, in_len, tl) for enc in encoder)
    loss_ce = sum(ce(dec.permute(0, 2, 1), trg_expect) for dec in att_dec)
    loss = 0.7 * loss_ce + 0.3 * loss_ctc
    m.zero_grad()
    loss.backward()
    assert abs(float(loss_ctc.detach()) - float(g["loss_ctc"])) < tol * float(g["loss_ctc"])
    assert abs(float(loss_ce.detach()) - float(g["loss_ce"])) < tol * float(g["loss_ce"])
    gtol = 5e-3 if precision == "fp32" else 6e-2
    P = dict(m.named_parameters())
    gmax = float(np.max(g["grad_norms"]))
    for name, ref_norm in zip(g["grad_names"], g["grad_norms"]):
        got = float(P[str(name)].grad.double().norm())
        assert abs(got - ref_norm) <= gtol * max(ref_norm, 1e-3 * gmax), (str(name), got, ref_norm)
    for k in [f[6:] for f in g.files if f.startswith("grad::")]:
        if k in P:
            assert rel(P[k].grad, torch.from_numpy(g["grad::" + k])) < gtol * 4, k
    # the fused loss kernel (eec_cross_entropy) == nn.CrossEntropyLoss on the same logits, value and gradient
    logits = att_dec.detach()[0].reshape(-1, 256).contiguous()
    tg = trg_expect.reshape(-1).contiguous()
    loss_k, dl = torch.zeros(1, device="cuda"), torch.empty_like(logits)
    eec.ops.cross_entropy(logits, tg, loss_k, dl)
    lr = logits.clone().requires_grad_(True)
    ref_l = ce(lr, tg)
    ref_l.backward()
    assert abs(float(loss_k) - float(ref_l)) < 1e-5 * abs(float(ref_l)) and rel(dl, lr.grad) < 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_long_utterance_max_len_vs_live_oracle(precision):
    """BASELINE configs[4] shape edge: T_in = 8001 -> T' = 1999 frames (max_len 2000, the positional table's limit), ragged
    lengths, one layer: log-probs and the multi-exit CTC loss against the CPU oracle run live."""
    import eec
    sd = O.make_params(21, n_exits=1, n_layers=1)
    src, lengths = O.synthetic_batch(2, 8001, seed=22)
    targets, tl = O.synthetic_targets(2, seed=23, lo=20, hi=60)
    m = eec.Early_conformer(src_pad_idx=0, n_enc_exits=1, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
                            d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.0, depthwise_kernel_size=31,
                            device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda").train()
    m.precision = precision
    out = m(src.cuda(), lengths)
    assert out.shape == (1, 2, 1999, 256)
    loss = eec.multi_exit_ctc_loss(out, targets, tl)
    loss.backward()
    ref = O.early_conformer_forward(sd, src, lengths, training=True)
    ref_loss, _, _ = O.multi_exit_ctc(ref.double(), targets, tl)
    assert rel(out.detach(), ref.detach()) < TOL[precision]
    assert abs(float(loss.detach()) - float(ref_loss)) < TOL[precision] * abs(float(ref_loss))
    assert float(m.conformer[0].conformer_layers[0].ffn1.sequential[1].weight.grad.norm()) > 0


def _noam_rate(step, d_model, warmup):   # util/noam_opt.py:35-40
    return d_model ** (-0.5) * min(step ** (-0.5), step * warmup ** (-1.5))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_noam_adamw_matches_torch(precision):
    """eec.FusedNoamAdamW == clip_grad_norm_ (train.py:69) + NoamOpt (util/noam_opt.py) + torch.optim.AdamW
    (train.py:261-262) applied on the CPU to the SAME gradients, over several steps; the bf16 operand shadow follows the
    weights (the next forward equals a forward of a fresh model holding the updated parameters)."""
    import eec
    name = "ec_e2l1_b3_t163"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, precision)
    m.train()
    targets, tl = O.synthetic_targets(src.shape[0], seed=31, lo=3, hi=6)
    warmup, clip, wd = 4, 0.05, 5e-4     # tiny warm-up so the rate moves; clip small enough to be active
    opt = eec.FusedNoamAdamW(m, model_size=256, warmup=warmup, betas=(0.9, 0.98), eps=1e-9, weight_decay=wd, clip=clip)
    ref_p = [p.detach().cpu().clone().requires_grad_(True) for p in m.parameters()]
    ref_opt = torch.optim.AdamW(ref_p, lr=0.0, betas=(0.9, 0.98), eps=1e-9, weight_decay=wd)
    x = src.cuda()
    for step in range(1, 4):
        opt.zero_grad()
        loss = eec.multi_exit_ctc_loss(m(x, lengths), targets, tl)
        loss.backward()
        grads = [p.grad.detach().cpu().clone() for p in m.parameters()]
        opt.step()
        for rp, gr in zip(ref_p, grads):
            rp.grad = gr
        total = torch.nn.utils.clip_grad_norm_(ref_p, clip)
        for grp in ref_opt.param_groups:
            grp["lr"] = _noam_rate(step, 256, warmup)
        ref_opt.step()
        assert opt._step == step
        assert abs(opt.last_grad_norm() - float(total)) < 1e-4 * float(total)
        assert abs(opt.rate() - _noam_rate(step, 256, warmup)) < 1e-12
        for (n, p), rp in zip(m.named_parameters(), ref_p):
            err = float((p.detach().cpu() - rp.detach()).abs().max())
            assert err <= 2e-6 * max(1.0, float(rp.detach().abs().max())), (step, n, err)
    # the next forward sees the updated weights (bf16: through the operand shadow written by the update kernel)
    m.eval()
    with torch.no_grad():
        out = m(x, lengths)
    m2, _, _, _ = build(name, g, precision)
    m2.load_state_dict(m.state_dict(), strict=True)
    m2.eval()
    with torch.no_grad():
        out2 = m2(x, lengths)
    assert rel(out, out2) < 1e-6


def test_graphed_step_with_fused_optimizer_trains():
    """One CUDA graph = forward + 6-exit CTC + backward + clip + Noam/AdamW; replaying it on a fixed batch lowers the loss
    and matches the eager sequence step for step."""
    import eec
    name = "ec_e2l1_b3_t163"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    targets, tl = O.synthetic_targets(3, seed=41, lo=3, hi=6)
    losses = {}
    for mode in ("eager", "graph"):
        m, sd, src, lengths = build(name, g, "bf16")
        m.train()
        opt = eec.FusedNoamAdamW(m, model_size=256, warmup=50, clip=1.0)
        x = src.cuda()
        if mode == "graph":
            step = eec.GraphedTrainStep(m, 3, src.shape[2], targets.shape[1], optimizer=opt)
            step.load_inputs(x, lengths, targets, tl)
            losses[mode] = [float(step.replay().clone()) for _ in range(6)]
            assert opt._step == 6
        else:
            ls = []
            for _ in range(6):
                opt.zero_grad()
                loss = eec.multi_exit_ctc_loss(m(x, lengths), targets, tl)
                loss.backward()
                opt.step()
                ls.append(float(loss.detach()))
            losses[mode] = ls
    assert losses["graph"][-1] < losses["graph"][0]
    for a, b in zip(losses["eager"], losses["graph"]):
        assert abs(a - b) < 2e-2 * abs(a), (losses["eager"], losses["graph"])


def test_graphed_forward_matches_eager():
    """eec.GraphedForward (inference forward as one CUDA graph, optionally truncated after exit e) == the eager forward."""
    import eec
    name = "ec_e3l2_b4_t331"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, "bf16")
    m.eval()
    with torch.no_grad():
        ref = m(src.cuda(), lengths)
    fwd = eec.GraphedForward(m, src.shape[0], src.shape[2])
    out = fwd(src, lengths)
    assert fwd.launches > 30 and rel(out, ref) < 1e-6
    src2, lengths2 = O.synthetic_batch(src.shape[0], src.shape[2], seed=5)
    with torch.no_grad():
        ref2 = m(src2.cuda(), lengths2)
    assert rel(fwd(src2.cuda(), lengths2), ref2) < 1e-6
    fwd1 = eec.GraphedForward(m, src.shape[0], src.shape[2], n_exits=1)
    assert rel(fwd1(src2, lengths2)[0], ref2[0]) < 1e-6 and fwd1.out.shape[0] == 1
    with pytest.raises(AssertionError):
        fwd(src2, torch.full((src.shape[0],), 100, dtype=torch.int64))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("split", [False, True])
def test_early_exit_compaction_skips_finished_rows(split, precision):
    """BASELINE configs[3]: (Splitformer) inference with dynamic early exit and on-device batch compaction.  After an exit the
    survivors are compacted and the later layers run on them only (eec_set_active_items: the kernels read the device-side
    count); decisions, tokens and the survivors' entropies must equal the selection rule applied to the FULL forward's output,
    for every threshold, eagerly and through the CUDA graph (one capture, thresholds/inputs change between replays)."""
    import eec
    E, Bn = 3, 7
    sd = O.make_params(13, n_exits=E, n_layers=1, splitformer=split)
    src, lengths = O.synthetic_batch(Bn, 331 if not split else 203, seed=14, min_frac=0.4)   # (203 -> odd T': the branch pads)
    cls = eec.Splitformer if split else eec.Early_conformer
    m = cls(src_pad_idx=0, n_enc_exits=E, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
            d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.1, depthwise_kernel_size=31,
            device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = precision
    with torch.no_grad():
        full = m(src.cuda(), lengths).cpu()
    T = full.shape[2]
    key_len = O.encoder_lengths(lengths, T)
    _, _, Hm = O.early_exit_select(full, key_len, 1e9)
    # thresholds: nobody exits early / a mix at every exit / everybody at the first exit
    srt = Hm[0].sort().values
    mids = [float((srt[2] + srt[3]) / 2), float(Hm[:2].max()) + 1e-3]
    for thr in [-1.0] + mids + [1e9]:
        ex_ref, tok_ref, _ = O.early_exit_select(full, key_len, thr)
        exit_index, tokens, n_tokens, mean_ent = m.forward_early_exit(src.cuda(), lengths, thr)
        ent_tol = 1e-3 if precision == "fp32" else 3e-2
        # a decision may legitimately flip only if an entropy sits within the arithmetic tolerance of the threshold
        margin = float((Hm - thr).abs().min())
        if margin > 2 * ent_tol:
            assert exit_index.cpu().tolist() == ex_ref.tolist(), (thr, exit_index.cpu().tolist(), ex_ref.tolist())
            if precision == "fp32":
                for b in range(Bn):
                    assert tokens[b, : int(n_tokens[b])].cpu().tolist() == tok_ref[b]
            for b in range(Bn):
                for e in range(int(ex_ref[b]) + 1):
                    assert abs(float(mean_ent[e, b]) - float(Hm[e, b])) < ent_tol, (thr, b, e)
    # graph: captured once, replayed on another batch
    thr = mids[0]
    ee = eec.GraphedEarlyExit(m, Bn, src.shape[2], thr)
    a = [t.clone() for t in ee(src, lengths)]
    b = m.forward_early_exit(src.cuda(), lengths, thr)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    src2, lengths2 = O.synthetic_batch(Bn, src.shape[2], seed=15, min_frac=0.4)
    a2 = [t.clone() for t in ee(src2, lengths2)]
    b2 = m.forward_early_exit(src2.cuda(), lengths2, thr)
    assert torch.equal(a2[0], b2[0]) and torch.equal(a2[1], b2[1])
    assert ee.launches > 20


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_edge_shapes_one_model_many_batches(precision):
    """The reference feeds length-sorted sub-batches of different shapes through one model (data_loader.py:166-188, train.py:26):
    the same eec model must take a stream of different (B, T_in) -- single utterance, T' = 1 and 2, ragged tile edges, T' just above
    one / two attention tiles -- in eval and in train mode, each checked against the CPU oracle run live."""
    import eec
    sd = O.make_params(23, n_exits=2, n_layers=1)
    m = eec.Early_conformer(src_pad_idx=0, n_enc_exits=2, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8,
                            max_len=2000, d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.0,
                            depthwise_kernel_size=31, device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    m.precision = precision
    shapes = [(1, 7), (1, 11), (2, 19), (5, 131), (1, 523), (3, 1031), (4, 67)]      # T' = 1, 2, 4, 32, 130, 257, 16
    for i, (Bn, t_in) in enumerate(shapes):
        src, lengths = O.synthetic_batch(Bn, t_in, seed=50 + i, min_frac=0.6)
        lengths[0] = t_in
        m.eval()
        with torch.no_grad():
            out = m(src.cuda(), lengths)
            ref = O.early_conformer_forward(sd, src, lengths)
        assert out.shape == ref.shape
        for e in range(2):
            assert rel(out[e], ref[e]) < TOL[precision], ("eval", Bn, t_in, e, rel(out[e], ref[e]))
        if Bn * out.shape[2] < 2:
            continue        # train-mode BatchNorm needs more than one value per channel (torch raises there too)
        m.train()
        running = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
        targets, tl = O.synthetic_targets(Bn, seed=70 + i, lo=1, hi=max(1, min(3, out.shape[2] // 2 - 1)))
        m.zero_grad(set_to_none=True)
        out_t = m(src.cuda(), lengths)
        loss = eec.multi_exit_ctc_loss(out_t, targets, tl)
        loss.backward()
        sdg = {k: (v.clone().double().requires_grad_(True) if v.is_floating_point() and "running" not in k and k != "positional_encoder.pe"
                   else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
        ref_t = O.early_conformer_forward(sdg, src.double(), lengths, training=True)
        in_len = torch.full((Bn,), ref_t.shape[2], dtype=torch.long)
        ref_loss = sum(torch.nn.functional.ctc_loss(ref_t[e].permute(1, 0, 2), targets, in_len, tl, blank=0, zero_infinity=True)
                       for e in range(2))
        for e in range(2):
            assert rel(out_t[e], ref_t[e].detach()) < TOL[precision], ("train", Bn, t_in, e)
        if float(ref_loss) > 0:
            assert abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()) < TOL[precision], (Bn, t_in, loss.item(), ref_loss.item())
            ref_loss.backward()
            gt = 5e-3 if precision == "fp32" else 6e-2
            gmax = max(float(sdg[n].grad.norm()) for n, _ in m.named_parameters())
            for n, p in m.named_parameters():
                err = abs(p.grad.double().norm().item() - sdg[n].grad.norm().item()) / max(sdg[n].grad.norm().item(), 1e-3 * gmax)
                assert err < gt, (Bn, t_in, n, err)
        m.load_state_dict({**m.state_dict(), **running})     # keep the BatchNorm running stats of `sd` for the next eval comparison


def test_optimizer_state_interchanges_with_reference_noamopt():
    """FusedNoamAdamW.state_dict() is a superset of the reference's NoamOpt.state_dict() (util/noam_opt.py:12-17:
    `_step`, `warmup`, `model_size`, `_rate`), and load_state_dict() accepts a reference `lr###-transformer` state (no Adam moments:
    they restart at zero, as in the reference, train.py:122-125 never saves them)."""
    import eec
    name = "ec_e2l1_b3_t163"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, "bf16")
    m.train()
    targets, tl = O.synthetic_targets(src.shape[0], seed=31, lo=3, hi=6)
    opt = eec.FusedNoamAdamW(m, model_size=256, warmup=7)
    for _ in range(2):
        opt.zero_grad()
        eec.multi_exit_ctc_loss(m(src.cuda(), lengths), targets, tl).backward()
        opt.step()
    st = opt.state_dict()
    assert st["_step"] == 2 and st["warmup"] == 7 and st["model_size"] == 256
    assert abs(st["_rate"] - _noam_rate(2, 256, 7)) < 1e-12

    class NoamOpt:   # the reference's scheduler restated (its state handling is `__dict__` based, util/noam_opt.py:12-25)
        def __init__(self):
            self.optimizer, self._step, self.warmup, self.model_size, self._rate = None, 0, 25000, 256, 0

        def state_dict(self):
            return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

        def load_state_dict(self, s):
            self.__dict__.update(s)
    ref = NoamOpt()
    ref.load_state_dict(st)                       # a file written here resumes the reference's schedule at the right step
    assert ref._step == 2 and ref.warmup == 7
    ref = NoamOpt()                               # a state the REFERENCE wrote: schedule only, no Adam moments
    ref._step, ref.warmup = 11, 13
    assert set(ref.state_dict()) == {"_step", "warmup", "model_size", "_rate"}
    opt.load_state_dict(ref.state_dict())         # ... and the reference's file loads here
    assert opt._step == 11 and opt.warmup == 13.0
    assert float(opt.exp_avg.abs().max()) == 0.0 and float(opt.exp_avg_sq.abs().max()) == 0.0
    opt.load_state_dict(st)
    assert opt._step == 2 and opt.warmup == 7.0 and float(opt.exp_avg.abs().max()) > 0.0


def test_truncated_splitformer_graph_keeps_first_branch():
    """GraphedForward(n_exits < E) on a Splitformer must equal the first exits of the FULL model (the parallel branch of
    group 0 is applied whatever follows, early_exit.py:314-356)."""
    import eec
    g = np.load(os.path.join(GOLDEN, "sf_e3l1_b2_t203.npz"))
    m, sd, src, lengths = build("sf_e3l1_b2_t203", g, "bf16")
    m.eval()
    with torch.no_grad():
        ref = m(src.cuda(), lengths)
    for n in (1, 2):
        fwd = eec.GraphedForward(m, src.shape[0], src.shape[2], n_exits=n)
        out = fwd(src, lengths)
        assert out.shape[0] == n
        for e in range(n):
            assert rel(out[e], ref[e]) < 1e-6, (n, e)


def test_graph_warmup_leaves_model_state_untouched():
    """building a GraphedTrainStep must not move BatchNorm running statistics, num_batches_tracked or the weights."""
    import eec
    name = "ec_e2l1_b3_t163"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd, src, lengths = build(name, g, "bf16")
    m.train()
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    step = eec.GraphedTrainStep(m, src.shape[0], src.shape[2], 12)
    torch.cuda.synchronize()
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    # two loads back to back without a sync in between: the second must not tear the first one's staged batch (pinned buffers are single)
    targets, tl = O.synthetic_targets(src.shape[0], seed=11, lo=3, hi=6)
    src2, lengths2 = O.synthetic_batch(src.shape[0], src.shape[2], seed=78)
    l1 = float(step(src, lengths, targets, tl).clone())
    step.load_inputs(src, lengths, targets, tl)
    a = step.replay().clone()
    step.load_inputs(src2, lengths2, targets, tl)
    b = step.replay().clone()
    torch.cuda.synchronize()
    assert abs(float(a) - l1) <= 1e-2 * abs(l1)
    step.load_inputs(src2, lengths2, targets, tl)
    b2 = step.replay().clone()
    torch.cuda.synchronize()
    assert torch.isfinite(a) and torch.isfinite(b) and torch.isfinite(b2)
