"""CPU restatement of the reference's feature extraction.  TEST INFRASTRUCTURE ONLY (same rules as conformer_oracle.py).

Reference: util/data_loader.py:7-18 -- ``torchaudio.transforms.Spectrogram(n_fft=args.n_fft*2, hop_length=args.hop_length,
win_length=args.win_length)`` then ``MelScale(sample_rate, n_mels, n_stft=args.n_fft+1)``, defaults util/conf.py:335-380
(16 kHz, n_fft 512 -> FFT 1024, win 320, hop 160, 80 mels).  torchaudio is a third-party, un-vendored dependency of the
reference; its arithmetic is restated here in numpy from the published definitions:
  torch.stft(center=True, pad_mode="reflect", window=hann_window(win, periodic) zero-padded to n_fft on both sides,
             onesided=True), power = 2;   melscale_fbanks(norm=None, mel_scale="htk").
Pinned by tests/golden/fbank_ref.npz = outputs of the real torchaudio transforms called exactly like data_loader.py:7-18
(oracle/make_golden.py::fbank_case)."""
from __future__ import annotations

import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=np.float32)
    m_min = 2595.0 * np.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * np.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2, dtype=np.float32)
    f_pts = (700.0 * (10 ** (m_pts / 2595.0) - 1.0)).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up)).astype(np.float32)          # (n_freqs, n_mels)


def fbank(wave: np.ndarray, sample_rate=16000, n_fft=512, win_length=320, hop_length=160, n_mels=80) -> np.ndarray:
    """wave (L,) float -> (n_mels, 1 + L // hop) power-mel features of ONE utterance (data_loader.py:124-125)."""
    fft = 2 * n_fft
    x = np.asarray(wave, dtype=np.float64)
    L = x.shape[0]
    pad = fft // 2
    xp = np.pad(x, (pad, pad), mode="reflect")
    T = 1 + L // hop_length
    w = np.zeros(fft)
    off = (fft - win_length) // 2
    w[off:off + win_length] = hann_periodic(win_length)
    frames = np.stack([xp[t * hop_length: t * hop_length + fft] * w for t in range(T)])     # (T, fft)
    spec = np.abs(np.fft.rfft(frames, n=fft, axis=1)) ** 2                                      # (T, fft/2 + 1)
    fb = melscale_fbanks_htk(fft // 2 + 1, 0.0, float(sample_rate // 2), n_mels, sample_rate).astype(np.float64)
    return (spec @ fb).T.astype(np.float32)


def fbank_batch(waves: np.ndarray, lengths, **kw):
    """padded batch (B, Lmax) + sample counts -> ((B, n_mels, Tmax) zero padded, frame counts) -- pad_sequence(..., 0), data_loader.py:21-26"""
    hop = kw.get("hop_length", 160)
    n_mels = kw.get("n_mels", 80)
    T = 1 + waves.shape[1] // hop
    out = np.zeros((waves.shape[0], n_mels, T), dtype=np.float32)
    frames = []
    for b, n in enumerate(lengths):
        f = fbank(waves[b, : int(n)], **kw)
        out[b, :, : f.shape[1]] = f
        frames.append(f.shape[1])
    return out, np.array(frames, dtype=np.int64)
