"""Generate tests/golden/*.npz by running the REAL reference (imported from
/root/reference, with torchaudio's Conformer and torch.nn.CTCLoss) on seeded inputs.

Run in the authoring container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures hold only inputs' seeds/shapes and the reference's OUTPUTS; parameters
are regenerated at test time by ``oracle.conformer_oracle.make_params(seed, ...)``
(the script proves they load into the reference module with ``strict=True``).
TEST INFRASTRUCTURE ONLY -- nothing under early-exit-transformer_b200/ imports this.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from models.model.early_exit import Early_conformer, Splitformer, full_conformer  # noqa: E402  (the reference)
from oracle import conformer_oracle as O  # noqa: E402

CASES = {
    # name: (model, n_exits, n_layers, B, T_in, target lo/hi)
    "ec_e2l1_b3_t163": ("early_conformer", 2, 1, 3, 163, 3, 8),
    "ec_e3l2_b4_t331": ("early_conformer", 3, 2, 4, 331, 4, 14),
    "sf_e2l1_b3_t166": ("splitformer", 2, 1, 3, 166, 3, 8),
    "sf_e3l1_b2_t203": ("splitformer", 3, 1, 2, 203, 3, 8),   # odd T' -> pad branch
}


def build_reference(kind, n_exits, n_layers, sd):
    cls = Early_conformer if kind == "early_conformer" else Splitformer
    m = cls(src_pad_idx=0, n_enc_exits=n_exits, enc_voc_size=256, dec_voc_size=256, d_model=256,
            n_head=8, max_len=2000, d_feed_forward=2048, n_enc_layers=n_layers, features_length=80,
            drop_prob=0.0, depthwise_kernel_size=31, device=torch.device("cpu"))
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def greedy_reference(emission):
    """util/beam_infer.py:21-23 restated with the same torch calls (the file itself cannot be
    imported here: flashlight-text is missing)."""
    idx = torch.argmax(emission, dim=-1)
    idx = torch.unique_consecutive(idx, dim=-1)
    return [int(i) for i in idx if i != 0]


def run_case(name, kind, n_exits, n_layers, B, t_in, lo, hi, seed):
    sd = O.make_params(seed, n_exits=n_exits, n_layers=n_layers, splitformer=(kind == "splitformer"))
    src, lengths = O.synthetic_batch(B, t_in, seed=seed + 1)
    targets, tl = O.synthetic_targets(B, seed=seed + 2, lo=lo, hi=hi)
    out = {"seed": seed, "B": B, "t_in": t_in, "n_exits": n_exits, "n_layers": n_layers,
           "lengths": lengths.numpy(), "targets": targets.numpy(), "target_lengths": tl.numpy()}

    # eval forward
    m = build_reference(kind, n_exits, n_layers, sd).eval()
    with torch.no_grad():
        lp = m(src, lengths)
    out["eval_logprobs"] = lp.numpy()
    E, _, T, _ = lp.shape
    toks = [[greedy_reference(lp[e, b]) for b in range(B)] for e in range(E)]
    flat = [t for e in toks for b in e for t in b]
    out["greedy_flat"] = np.array(flat, dtype=np.int64)
    out["greedy_counts"] = np.array([[len(b) for b in e] for e in toks], dtype=np.int64)

    # train-mode forward + multi-exit CTC + backward (train.py:53-68, drop_prob=0)
    m = build_reference(kind, n_exits, n_layers, sd).train()
    lp = m(src, lengths)
    ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    in_len = torch.full((B,), lp.size(2), dtype=torch.long)
    per = [ctc(enc.permute(1, 0, 2), targets, in_len, tl) for enc in lp]
    loss = sum(per)
    m.zero_grad()
    loss.backward()
    out["train_logprobs"] = lp.detach().numpy()
    out["loss"] = np.float64(loss.item())
    out["loss_per_exit"] = np.array([p.item() for p in per])
    names, norms = [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(p.grad.double().norm().item())
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    # a few full gradients (small tensors + slices of the big ones)
    P = dict(m.named_parameters())
    l0 = "conformer.0.conformer_layers.0."
    for k in ["conv_subsample.sequential.0.bias", "conv_subsample.sequential.1.bias", "linears.0.bias",
              l0 + "ffn1.sequential.0.weight", l0 + "ffn1.sequential.1.bias", l0 + "self_attn.in_proj_bias",
              l0 + "conv_module.sequential.2.weight", l0 + "conv_module.sequential.3.weight",
              l0 + "conv_module.sequential.3.bias", l0 + "final_layer_norm.weight",
              l0 + "conv_module.sequential.0.bias", l0 + "self_attn.out_proj.bias"]:
        out["grad::" + k] = P[k].grad.numpy()
    out["gradslice::" + l0 + "ffn1.sequential.1.weight"] = P[l0 + "ffn1.sequential.1.weight"].grad[:8, :].numpy()
    out["gradslice::" + l0 + "self_attn.in_proj_weight"] = P[l0 + "self_attn.in_proj_weight"].grad[::96, :].numpy()
    out["gradslice::conv_subsample.sequential.0.weight"] = P["conv_subsample.sequential.0.weight"].grad[:4].numpy()
    out["gradslice::linears.0.weight"] = P["linears.0.weight"].grad[:8].numpy()
    bn = l0 + "conv_module.sequential.3."
    msd = m.state_dict()
    out["bn_running_mean"] = msd[bn + "running_mean"].numpy()
    out["bn_running_var"] = msd[bn + "running_var"].numpy()
    out["bn_num_batches_tracked"] = msd[bn + "num_batches_tracked"].numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "T'=", T, "loss=", loss.item(), os.path.getsize(path) // 1024, "KiB")


# --------------------------------------------------------------------------- #
# train mode WITH dropout (the reference's default --drop_prob 0.1, util/conf.py:283-291)
# --------------------------------------------------------------------------- #
DROP_CASES = {
    # name: (model, n_exits, n_layers, B, T_in, target lo/hi, p, dropout seed)
    "ecdrop_e2l2_b3_t203": ("early_conformer", 2, 2, 3, 203, 3, 8, 0.1, 31337),
    "sfdrop_e2l1_b3_t166": ("splitformer", 2, 1, 3, 166, 3, 8, 0.1, 4242),
}


class ReferenceDropoutPatch:
    """Runs the UNMODIFIED reference with its random draws replaced by given masks.  torch's RNG streams cannot be matched by
    another implementation, so the parity question for dropout is: given the same keep-masks at the same places, does the
    product compute what the reference computes?  Every mask the reference consumes goes through two functions:
    ``torch.nn.functional.dropout`` (all nn.Dropout modules) and ``torch.nn.functional.scaled_dot_product_attention`` (the
    dropout_p of nn.MultiheadAttention, need_weights=False branch, TORCH:nn/functional.py:6670-6695).  Both are patched
    for the duration of one forward; the k-th call is mapped to the product's documented site id and the mask
    (oracle/philox_dropout.Masks, element order (b, t, channel)) is permuted into the layout the reference holds there:
      site 0            positional_encoding.py:72        (B, T, D)
      layer sites 0,1   TA:106, TA:108 (ffn1)            (T, B, F) / (T, B, D)
      layer site  2     SDPA probabilities (TA:152)      (B, H, T, T), rows of stride 8*ceil(T/8)
      layer site  3     TA:201 self_attn_dropout         (T, B, D)
      layer site  4     TA:73 conv-module dropout        (B, D, T)
      layer sites 5,6   TA:106, TA:108 (ffn2)            (T, B, F) / (T, B, D)
    """

    def __init__(self, masks, layer_uids):
        self.masks, self.uids = masks, list(layer_uids)
        self.calls = 0

    def _site(self):
        k = self.calls
        self.calls += 1
        if k == 0:
            return 0, None
        layer, j = divmod(k - 1, 7)
        return 8 * (self.uids[layer] + 1) + j, j

    def dropout(self, input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        site, j = self._site()
        assert j != 2, "call order drifted: expected the SDPA call here"
        if j is None:
            f = torch.as_tensor(self.masks(site, tuple(input.shape)))
        elif j == 4:
            B, D, T = input.shape
            f = torch.as_tensor(self.masks(site, (B, T, D))).permute(0, 2, 1)
        else:
            T, B, C = input.shape
            f = torch.as_tensor(self.masks(site, (B, T, C))).permute(1, 0, 2)
        return input * f.to(input.dtype)

    def sdpa(self, query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, **kw):
        import math
        assert not is_causal
        s = query @ key.transpose(-1, -2) / math.sqrt(query.shape[-1])
        if attn_mask is not None:
            s = s.masked_fill(attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else s + attn_mask
        pr = torch.nan_to_num(torch.softmax(s, -1), nan=0.0)   # fully masked rows -> 0 (torch 2.11 CPU SDPA)
        if dropout_p > 0.0:
            site, j = self._site()
            assert j == 2, "call order drifted: SDPA expected at layer slot 2"
            T = pr.shape[-1]
            pr = pr * torch.as_tensor(self.masks(site, tuple(pr.shape), 8 * ((T + 7) // 8))).to(pr.dtype)
        return pr @ value

    def __enter__(self):
        import torch.nn.functional as F
        self._saved = (F.dropout, F.scaled_dot_product_attention)
        F.dropout, F.scaled_dot_product_attention = self.dropout, self.sdpa
        return self

    def __exit__(self, *a):
        import torch.nn.functional as F
        F.dropout, F.scaled_dot_product_attention = self._saved


def run_dropout_case(name, kind, n_exits, n_layers, B, t_in, lo, hi, p, dseed, seed):
    from oracle import philox_dropout as PH
    split = kind == "splitformer"
    sd = O.make_params(seed, n_exits=n_exits, n_layers=n_layers, splitformer=split)
    src, lengths = O.synthetic_batch(B, t_in, seed=seed + 1)
    targets, tl = O.synthetic_targets(B, seed=seed + 2, lo=lo, hi=hi)
    cls = Splitformer if split else Early_conformer
    m = cls(src_pad_idx=0, n_enc_exits=n_exits, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
            d_feed_forward=2048, n_enc_layers=n_layers, features_length=80, drop_prob=p, depthwise_kernel_size=31,
            device=torch.device("cpu"))
    m.load_state_dict(sd, strict=True)
    m.train()
    uids = []      # layer execution order of the reference forward (early_exit.py:626-631 / :305-356)
    for e in range(n_exits):
        uids += [e * n_layers + l for l in range(n_layers)]
        if split and e in (0, n_exits - 1):
            uids.append(1000 + e // (n_exits - 1))
    # sanity of the patch itself: with p-effective 0 masks (all ones) the patched forward equals the unpatched p = 0 model
    patch = ReferenceDropoutPatch(PH.Masks(p, dseed, 0), uids)
    with patch:
        lp = m(src, lengths)
    assert patch.calls == 1 + 7 * len(uids), (patch.calls, len(uids))
    ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    in_len = torch.full((B,), lp.size(2), dtype=torch.long)
    per = [ctc(enc.permute(1, 0, 2), targets, in_len, tl) for enc in lp]
    loss = sum(per)
    m.zero_grad()
    loss.backward()
    out = {"seed": seed, "B": B, "t_in": t_in, "n_exits": n_exits, "n_layers": n_layers, "p": p, "drop_seed": dseed,
           "lengths": lengths.numpy(), "targets": targets.numpy(), "target_lengths": tl.numpy(),
           "train_logprobs": lp.detach().numpy(), "loss": np.float64(loss.item()),
           "loss_per_exit": np.array([q.item() for q in per])}
    names, norms = [], []
    for k, q in m.named_parameters():
        names.append(k)
        norms.append(q.grad.double().norm().item())
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms)
    P = dict(m.named_parameters())
    l0 = "conformer.0.conformer_layers.0."
    for k in ["conv_subsample.sequential.1.bias", "linears.0.bias", l0 + "ffn1.sequential.1.bias", l0 + "ffn1.sequential.4.bias",
              l0 + "self_attn.in_proj_bias", l0 + "self_attn.out_proj.bias", l0 + "conv_module.sequential.5.bias",
              l0 + "conv_module.sequential.2.weight", l0 + "ffn2.sequential.4.bias", l0 + "final_layer_norm.weight"]:
        out["grad::" + k] = P[k].grad.numpy()
    out["gradslice::" + l0 + "ffn1.sequential.1.weight"] = P[l0 + "ffn1.sequential.1.weight"].grad[:8, :].numpy()
    out["gradslice::" + l0 + "self_attn.in_proj_weight"] = P[l0 + "self_attn.in_proj_weight"].grad[::96, :].numpy()
    out["gradslice::conv_subsample.sequential.0.weight"] = P["conv_subsample.sequential.0.weight"].grad[:4].numpy()
    # the patch is transparent: same patched functions, p-independent all-ones masks == the unpatched reference at p = 0
    m0 = build_reference(kind, n_exits, n_layers, sd).train()
    m0p = cls(src_pad_idx=0, n_enc_exits=n_exits, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
              d_feed_forward=2048, n_enc_layers=n_layers, features_length=80, drop_prob=p, depthwise_kernel_size=31,
              device=torch.device("cpu"))
    m0p.load_state_dict(sd, strict=True)
    m0p.train()
    with torch.no_grad():
        a = m0(src, lengths)
        with ReferenceDropoutPatch(lambda site, shape, row_stride=None: np.ones(shape, dtype=np.float32), uids):
            b = m0p(src, lengths)
    assert float((a - b).abs().max()) < 2e-5, float((a - b).abs().max())
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "T'=", lp.shape[2], "loss=", loss.item(), "(p=0 loss differs)", os.path.getsize(path) // 1024, "KiB")


def fbank_case():
    """tests/golden/fbank_ref.npz: the real torchaudio transforms called exactly like util/data_loader.py:7-18 (the reference's own
    util.data_loader cannot be imported here -- no matter: these two calls ARE its feature extraction) on seeded waveforms."""
    import torchaudio.transforms as TT
    from types import SimpleNamespace
    args = SimpleNamespace(n_fft=512, hop_length=160, win_length=320, sample_rate=16000, n_mels=80)   # util/conf.py:335-380
    g = torch.Generator().manual_seed(77)
    lens = [16000, 9999, 4803, 1600]
    out = {"lengths": np.array(lens, dtype=np.int64), "seed": 77}
    for i, n in enumerate(lens):
        wave = torch.randn(1, n, generator=g) * 0.1
        wave[0, : n // 3] += 0.3 * torch.sin(torch.arange(n // 3) * 2 * math.pi * 440.0 / 16000)
        spec = TT.Spectrogram(n_fft=args.n_fft * 2, hop_length=args.hop_length, win_length=args.win_length)(wave)
        mel = TT.MelScale(sample_rate=args.sample_rate, n_mels=args.n_mels, n_stft=args.n_fft + 1)(spec)
        out[f"wave{i}"] = wave[0].numpy()
        out[f"fbank{i}"] = mel[0].numpy()
    path = os.path.join(ROOT, "tests", "golden", "fbank_ref.npz")
    np.savez_compressed(path, **out)
    print("fbank_ref", [out[f"fbank{i}"].shape for i in range(len(lens))], os.path.getsize(path) // 1024, "KiB")


FC_RENAME = (("linears.", "linears_1."), ("positional_encoder.", "positional_encoder_1."))
AED_CE_WEIGHT, AED_CTC_WEIGHT = 0.7, 0.3   # util/conf.py --aed_ce_weight / --aed_ctc_weight defaults


def fc_state_dict(sd):
    """Early_conformer-layout encoder parameters -> full_conformer key names (early_exit.py:671-686)."""
    out = {}
    for k, v in sd.items():
        for a, b in FC_RENAME:
            if k.startswith(a):
                k = b + k[len(a):]
                break
        out[k] = v
    return out


def run_fc_case(name, n_exits, n_layers, n_dec, B, t_in, lo, hi, seed):
    """full_conformer (AED mode, early_exit.py:637-811): encoder parameters from make_params, decoder /
    embedding parameters = the reference's own default init under torch.manual_seed(seed) (regenerated at test time)."""
    kw = dict(trg_pad_idx=126, n_enc_exits=n_exits, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
              d_feed_forward=2048, n_enc_layers=n_layers, n_dec_layers=n_dec, features_length=80, drop_prob=0.0,
              depthwise_kernel_size=31, device=torch.device("cpu"))
    enc_sd = fc_state_dict(O.make_params(seed, n_exits=n_exits, n_layers=n_layers))
    src, lengths = O.synthetic_batch(B, t_in, seed=seed + 1)
    targets, tl = O.synthetic_targets(B, seed=seed + 2, lo=lo, hi=hi)
    trg, trg_expect = targets[:, :-1], targets[:, 1:]     # train.py:30-32

    def build():
        torch.manual_seed(seed)
        m = full_conformer(**kw)
        r = m.load_state_dict(enc_sd, strict=False)
        assert not r.unexpected_keys and all(k.startswith(("decoders.", "emb.", "layer_norm.", "linears_2.", "positional_encoder_2."))
                                             for k in r.missing_keys), r
        return m

    out = {"seed": seed, "B": B, "t_in": t_in, "n_exits": n_exits, "n_layers": n_layers, "n_dec": n_dec,
           "lengths": lengths.numpy(), "targets": targets.numpy(), "target_lengths": tl.numpy()}
    m = build().eval()
    with torch.no_grad():
        dec_out, enc_out = m(src, lengths, trg)
        enc1 = m._encoder_(src, lengths, 1)
        dec1 = m._decoder_(trg, enc1, 1)
    out.update(eval_dec_out=dec_out.numpy(), eval_enc_out=enc_out.numpy(), encoder_1=enc1.numpy(), decoder_1=dec1.numpy())
    m = build().train()
    att_dec, encoder = m(src, lengths, trg)
    ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    ce = torch.nn.CrossEntropyLoss()
    in_len = torch.full((B,), encoder.size(2), dtype=torch.long)
    loss_ctc = sum(ctc(enc.permute(1, 0, 2), targets, in_len, tl) for enc in encoder)        # train.py:44-46
    loss_ce = sum(ce(dec.permute(0, 2, 1), trg_expect) for dec in att_dec)                   # train.py:47
    loss = AED_CE_WEIGHT * loss_ce + AED_CTC_WEIGHT * loss_ctc                                # train.py:51
    m.zero_grad()
    loss.backward()
    out["loss"], out["loss_ce"], out["loss_ctc"] = np.float64(loss.item()), np.float64(loss_ce.item()), np.float64(loss_ctc.item())
    names, norms = [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else p.grad.double().norm().item())
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms)
    P = dict(m.named_parameters())
    l0 = "conformer.0.conformer_layers.0."
    d0 = "decoders.0.layers.0."
    extra = [d0 + "self_attn.in_proj_bias", d0 + "multihead_attn.in_proj_weight", d0 + "linear1.bias", d0 + "norm2.weight", "emb.weight"] if n_dec > 1 else []
    for k in ["conv_subsample.sequential.0.bias", "linears_1.0.bias", "linears_2.1.bias", l0 + "ffn1.sequential.1.bias",
              l0 + "final_layer_norm.weight", "layer_norm.weight"] + extra:
        out["grad::" + k] = P[k].grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss=", loss.item(), "ce=", loss_ce.item(), "ctc=", loss_ctc.item(), os.path.getsize(path) // 1024, "KiB")


def ctc_cases():
    """Stand-alone CTC known answers from torch.nn.CTCLoss incl. repeats, empty and
    infeasible targets (zero_infinity)."""
    g = torch.Generator().manual_seed(99)
    T, B, V = 24, 6, 256
    lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * 2, -1).requires_grad_(True)
    targets = torch.randint(1, V, (B, 14), generator=g)
    targets[1, :6] = 7          # repeated labels
    tl = torch.tensor([5, 6, 14, 0, 1, 11])
    targets[2, :14] = 9         # 14 repeats need 27 frames > T -> infeasible
    ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)
    loss = ctc(lp.permute(1, 0, 2), targets, torch.full((B,), T, dtype=torch.long), tl)
    loss.backward()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ctc_kat.npz"), lp=lp.detach().numpy(),
                        targets=targets.numpy(), target_lengths=tl.numpy(), loss=np.float64(loss.item()),
                        grad=lp.grad.numpy())
    print("ctc_kat loss=", loss.item())


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "fc2":     # only the second AED fixture (two decoder layers per exit, ragged targets with padding)
        run_fc_case("fc_e2l1d2_b3_t203", 2, 1, 2, 3, 203, 3, 9, seed=160)
        sys.exit(0)
    for i, (name, (kind, e, l, B, t, lo, hi)) in enumerate(CASES.items()):
        run_case(name, kind, e, l, B, t, lo, hi, seed=100 + 10 * i)
    ctc_cases()
    run_fc_case("fc_e2l1d1_b2_t163", 2, 1, 1, 2, 163, 3, 8, seed=150)
    run_fc_case("fc_e2l1d2_b3_t203", 2, 1, 2, 3, 203, 3, 9, seed=160)
    fbank_case()
    for i, (name, (kind, e, l, B, t, lo, hi, p, ds)) in enumerate(DROP_CASES.items()):
        run_dropout_case(name, kind, e, l, B, t, lo, hi, p, ds, seed=200 + 10 * i)
