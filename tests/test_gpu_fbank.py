"""GPU parity of the feature front end (util/data_loader.py:7-18 = torchaudio Spectrogram + MelScale) against the
golden outputs of the real torchaudio transforms and the numpy oracle.  Tolerance: the north star's fp32 bar, 1e-3 relative
(maxabs(a-b)/maxabs(b)); the split-bf16 tensor-core GEMMs land around 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import conformer_oracle as O
from oracle import fbank_oracle as FO

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_fbank_vs_torchaudio_golden_ragged_batch():
    import eec
    g = np.load(os.path.join(GOLDEN, "fbank_ref.npz"))
    lens = [int(n) for n in g["lengths"]]
    waves = torch.zeros(len(lens), max(lens))
    for i, n in enumerate(lens):
        waves[i, :n] = torch.from_numpy(g[f"wave{i}"])
    fb = eec.Fbank()
    feats, frames = fb(waves.cuda(), torch.tensor(lens))
    assert feats.shape == (len(lens), 80, 1 + max(lens) // 160) and feats.dtype == torch.float32
    assert frames.cpu().tolist() == [1 + n // 160 for n in lens]
    feats = feats.cpu().numpy()
    for i, n in enumerate(lens):
        T = 1 + n // 160
        assert rel(feats[i, :, :T], g[f"fbank{i}"]) < 1e-3, (i, rel(feats[i, :, :T], g[f"fbank{i}"]))
        assert rel(feats[i, :, :T], g[f"fbank{i}"]) < 1e-4          # what the hi/lo split actually achieves
        assert float(np.abs(feats[i, :, T:]).max(initial=0.0)) == 0.0     # zero padding like pad_sequence(..., 0)
    with pytest.raises(eec.EecError):
        fb(waves, torch.tensor(lens))     # CPU tensor: no fallback


def test_fbank_full_size_vs_oracle_and_into_the_model():
    """BASELINE-size utterances (15 s) for a few rows against the numpy oracle, heavy-tailed amplitudes; then waveform -> fbank ->
    Early_conformer on the GPU equals the oracle features -> CPU oracle model."""
    import eec
    g = torch.Generator().manual_seed(5)
    Bn, L = 3, 240000
    waves = torch.randn(Bn, L, generator=g) * torch.tensor([0.01, 0.3, 2.0])[:, None]
    lens = [240000, 200001, 160159]
    for i, n in enumerate(lens):
        waves[i, n:] = 0
    fb = eec.Fbank()
    feats, frames = fb(waves.cuda(), torch.tensor(lens))
    ref, ref_frames = FO.fbank_batch(waves.numpy(), lens)
    assert frames.cpu().tolist() == ref_frames.tolist() and feats.shape == ref.shape == (Bn, 80, 1501)
    for i in range(Bn):
        assert rel(feats[i].cpu().numpy(), ref[i]) < 1e-4, (i, rel(feats[i].cpu().numpy(), ref[i]))
    sd = O.make_params(31, n_exits=1, n_layers=1)
    m = eec.Early_conformer(src_pad_idx=0, n_enc_exits=1, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
                            d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.0, depthwise_kernel_size=31,
                            device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(feats, frames.cpu())
        want = O.early_conformer_forward(sd, torch.from_numpy(ref), torch.from_numpy(ref_frames))
    assert float((out.cpu() - want).abs().max() / want.abs().max()) < 1e-3


def test_waveform_to_loss_pipeline_trains_with_dropout():
    """The whole replaced pipeline on a fixed batch: waveform -> eec.Fbank -> Early_conformer (train mode, the reference's default
    drop_prob 0.1) -> fused multi-exit CTC -> clip + AdamW, as ONE CUDA graph per step.  The loss must fall while dropout draws
    fresh masks every replay."""
    import eec
    g = torch.Generator().manual_seed(3)
    Bn, L = 4, 32000
    waves = torch.randn(Bn, L, generator=g) * 0.1
    lens = torch.tensor([32000, 30000, 28000, 32000])
    feats, frames = eec.Fbank()(waves.cuda(), lens)
    feats = torch.log1p(feats)                         # (tame the power-mel dynamic range for this tiny optimisation problem)
    targets, tl = O.synthetic_targets(Bn, seed=4, lo=3, hi=6)
    sd = O.make_params(41, n_exits=2, n_layers=1)
    m = eec.Early_conformer(src_pad_idx=0, n_enc_exits=2, enc_voc_size=256, dec_voc_size=256, d_model=256, n_head=8, max_len=2000,
                            d_feed_forward=2048, n_enc_layers=1, features_length=80, drop_prob=0.1, depthwise_kernel_size=31,
                            device=torch.device("cuda"))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    m.precision = "bf16"
    m.set_dropout_seed(7)
    opt = eec.FusedNoamAdamW(m, model_size=256, warmup=50, clip=1.0, lr=3e-4)
    step = eec.GraphedTrainStep(m, Bn, feats.shape[2], targets.shape[1], optimizer=opt)
    step.load_inputs(feats, frames.cpu(), targets, tl)
    losses = [float(step.replay().clone()) for _ in range(40)]
    assert all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < 0.9 * np.mean(losses[:5]), losses
    assert len(set(round(x, 4) for x in losses)) > 30          # fresh masks + moving weights: no two steps alike
