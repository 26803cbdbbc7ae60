"""Feature front end on the GPU: the reference's ``spec_transform`` + ``melspec_transform``
(util/data_loader.py:7-18: torchaudio ``Spectrogram(n_fft=2*args.n_fft, hop_length, win_length)`` -> ``MelScale(sample_rate,
n_mels, n_stft=args.n_fft+1)``, applied per utterance in the collate functions, :124-125, :200-201, :256-257) for a whole
padded batch of waveforms, through libeec.so (csrc/fbank.cu + the tcgen05 GEMM).

    fbank = eec.Fbank().cuda_tables("cuda")                       # defaults = util/conf.py: 16 kHz, n_fft 512 (FFT 1024), win 320, hop 160, 80 mels
    feats, lengths = fbank(waveforms, wave_lengths)               # (B, L) fp32 CUDA, (B,) int64  ->  (B, 80, T) fp32, (B,) int64 frames
    log_probs = model(feats, lengths.cpu())                       # the tensors data_loader.py hands to the model

Utterance b yields 1 + L_b // hop frames (torch.stft, center=True); frames past that are zero, exactly like the reference's
``pad_sequence(..., 0)`` of per-utterance features (data_loader.py:21-26).  There is no CPU path."""
from __future__ import annotations

import math

import torch

from . import ops
from .lib import EecError, call, ptr, stream


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") restated (the table MelScale holds as ``fb``): (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class Fbank:
    def __init__(self, sample_rate: int = 16000, n_fft: int = 512, win_length: int = 320, hop_length: int = 160, n_mels: int = 80):
        # (argument names and defaults follow util/conf.py:335-380; the reference's FFT size is 2 * args.n_fft, data_loader.py:8)
        self.sample_rate, self.fft, self.win, self.hop, self.n_mels = sample_rate, 2 * n_fft, win_length, hop_length, n_mels
        self.n_freqs = self.fft // 2 + 1
        if win_length % 64 or win_length > self.fft:
            raise EecError(f"Fbank: win_length {win_length} unsupported: need a multiple of 64 that fits the FFT size {self.fft} "
                           "(the reference uses 320)")
        self.n_spec = ((2 * self.n_freqs + 31) // 32) * 32          # [re | im] columns padded to the GEMM's N granularity
        self.kp = ((self.n_freqs + 63) // 64) * 64                    # power columns padded so that 3*kp is a multiple of the GEMM's k-block
        self.n_mel_pad = ((n_mels + 31) // 32) * 32
        self.dev = None

    def cuda_tables(self, device) -> "Fbank":
        """Build the constant tables on `device`: hann window, [cos | sin] DFT rows restricted to the window support, mel filterbank."""
        dev = torch.device(device)
        win, fft, nf = self.win, self.fft, self.n_freqs
        self.window = torch.hann_window(win, periodic=True).to(dev)
        off = (fft - win) // 2                                         # torch.stft centres a short window inside the FFT frame
        n = torch.arange(win, dtype=torch.float64) + off
        f = torch.arange(nf, dtype=torch.float64)
        ang = 2.0 * math.pi * f[:, None] * n[None, :] / fft
        dft = torch.zeros(self.n_spec, win, dtype=torch.float64)
        dft[:nf] = torch.cos(ang)
        dft[nf:2 * nf] = torch.sin(ang)                                # (the sign of the imaginary part does not survive |.|^2)
        fb = torch.zeros(self.n_mel_pad, nf)
        fb[: self.n_mels] = melscale_fbanks_htk(nf, 0.0, float(self.sample_rate // 2), self.n_mels, self.sample_rate).t()
        dft32, fb32 = dft.float().to(dev).contiguous(), fb.to(dev).contiguous()
        self.dft3 = torch.empty(self.n_spec, 3 * win, dtype=torch.bfloat16, device=dev)
        self.fb3 = torch.empty(self.n_mel_pad, 3 * self.kp, dtype=torch.bfloat16, device=dev)
        call("eec_fbank_split_operand", ptr(dft32), win, ptr(self.dft3), self.n_spec, win, stream())
        call("eec_fbank_split_operand", ptr(fb32), nf, ptr(self.fb3), self.n_mel_pad, self.kp, stream())
        self.dev = dev
        return self

    def n_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop

    def __call__(self, wave: torch.Tensor, wave_lengths: torch.Tensor):
        if not wave.is_cuda:
            raise EecError("Fbank: waveforms must be on a CUDA device; there is no CPU path")
        if self.dev is None or self.dev != wave.device:
            self.cuda_tables(wave.device)
        wave = wave.contiguous().float()
        B, L = wave.shape
        T = self.n_frames(L)
        dev, f32, bf16 = wave.device, torch.float32, torch.bfloat16
        wl = wave_lengths.to(device=dev, dtype=torch.int64, non_blocking=True)
        M = B * T
        frames = torch.empty(M, 3 * self.win, dtype=bf16, device=dev)
        call("eec_fbank_frames", ptr(wave), ptr(wl), L, ptr(self.window), ptr(frames), B, T, self.win, self.hop, stream())
        spec = torch.empty(M, self.n_spec, dtype=f32, device=dev)
        ops.gemm(frames, self.dft3, spec, M, self.n_spec, 3 * self.win)
        power = torch.empty(M, 3 * self.kp, dtype=bf16, device=dev)
        call("eec_fbank_power", ptr(spec), self.n_spec, ptr(power), M, self.n_freqs, self.kp, stream())
        mel = torch.empty(M, self.n_mel_pad, dtype=f32, device=dev)
        ops.gemm(power, self.fb3, mel, M, self.n_mel_pad, 3 * self.kp)
        out = torch.empty(B, self.n_mels, T, dtype=f32, device=dev)
        call("eec_fbank_finish", ptr(mel), self.n_mel_pad, ptr(out), B, T, self.n_mels, stream())
        lengths = 1 + torch.div(wl, self.hop, rounding_mode="floor")   # (plumbing: the frame counts handed to the model)
        return out, lengths
