"""Generate tests/golden/*.npz by running the REAL reference (imported from
/root/reference, with torchaudio's Conformer and torch.nn.CTCLoss) on seeded inputs.

Run in the authoring container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures hold only inputs' seeds/shapes and the reference's OUTPUTS; parameters
are regenerated at test time by ``oracle.conformer_oracle.make_params(seed, ...)``
(the script proves they load into the reference module with ``strict=True``).
TEST INFRASTRUCTURE ONLY -- nothing under early-exit-transformer_b200/ imports this.
"""
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
    for k in ["conv_subsample.sequential.0.bias", "linears_1.0.bias", "linears_2.1.bias", l0 + "ffn1.sequential.1.bias",
              l0 + "final_layer_norm.weight", "layer_norm.weight"]:
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
    for i, (name, (kind, e, l, B, t, lo, hi)) in enumerate(CASES.items()):
        run_case(name, kind, e, l, B, t, lo, hi, seed=100 + 10 * i)
    ctc_cases()
    run_fc_case("fc_e2l1d1_b2_t163", 2, 1, 1, 2, 163, 3, 8, seed=150)
