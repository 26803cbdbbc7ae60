"""CPU restatement of the counter-based dropout masks of libeec.so.  TEST INFRASTRUCTURE ONLY
(same rules as conformer_oracle.py: imported by tests/, smoke() and bench.py's CPU legs, never by the product).

The reference applies ``torch.nn.Dropout`` (positional_encoding.py:72; torchaudio conformer.py TA:73, TA:106, TA:108,
TA:201) and ``nn.MultiheadAttention(dropout=p)`` (TA:152) in train mode.  PyTorch's RNG streams cannot be reproduced by
another implementation (SURVEY App. A "Dropout sites"), so the product defines its own, documented in include/eec.h:

    element i of the logical tensor at dropout site s is KEPT iff
        u16_{i % 8}( Philox4x32-10( key = (seed_lo, seed_hi ^ offset_hi),
                                    counter = (i/8 lo, i/8 hi, s, offset_lo) ) ) >= thr,     thr = round(p * 65536)
    and kept values are scaled by 65536 / (65536 - thr).

Philox4x32-10 is the published generator of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11) -- the
same one torch's CUDA generator uses; `philox4x32_10` below is pinned by the Random123 known-answer vectors in
tests/test_oracle_golden.py.  With these masks injected into conformer_oracle.early_conformer_forward(drop=...) the
dropout path of the CUDA kernels is checked value-for-value, forward and backward.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key) -> np.ndarray:
    """ctr: uint32 [..., 4]; key: (k0, k1) python ints -> uint32 [..., 4].  Ten rounds, key bumped by the Weyl constants."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        n0 = (p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1)
        c = [n0 & MASK32, p1 & MASK32, n2 & MASK32, p0 & MASK32]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def threshold(p: float) -> int:
    return min(int(p * 65536.0 + 0.5), 65535)


def p_effective(p: float) -> float:
    return threshold(p) / 65536.0


def draws_u16(n: int, seed: int, offset: int, site: int) -> np.ndarray:
    """the 16-bit draws of elements 0..n-1 of one site (uint16 [n])"""
    groups = (n + 7) // 8
    g = np.arange(groups, dtype=np.uint64)
    ctr = np.empty((groups, 4), dtype=np.uint32)
    ctr[:, 0] = (g & MASK32).astype(np.uint32)
    ctr[:, 1] = (g >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(site & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(offset & 0xFFFFFFFF)
    key = (seed & 0xFFFFFFFF, ((seed >> 32) ^ (offset >> 32)) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key)                                   # [groups, 4] uint32
    u = np.empty((groups, 8), dtype=np.uint16)
    u[:, 0::2] = (r & np.uint32(0xFFFF)).astype(np.uint16)
    u[:, 1::2] = (r >> np.uint32(16)).astype(np.uint16)
    return u.reshape(-1)[:n]


def factors(n: int, p: float, seed: int, offset: int, site: int) -> np.ndarray:
    """float32 [n]: 65536/(65536-thr) where element i is kept, 0 where it is dropped"""
    thr = threshold(p)
    if thr == 0:
        return np.ones(n, dtype=np.float32)
    scale = np.float32(65536.0 / (65536.0 - thr))
    return np.where(draws_u16(n, seed, offset, site) >= thr, scale, np.float32(0.0)).astype(np.float32)


class Masks:
    """Callable handed to conformer_oracle.*(drop=...): drop(site, shape, row_stride=None) -> numpy float32 factors.
    `shape` is the logical tensor; when row_stride is given the last dimension is embedded in rows of that stride
    (the attention probabilities: stride 8*ceil(T/8), include/eec.h eec_attn_fwd)."""

    def __init__(self, p: float, seed: int, offset: int = 0):
        self.p, self.seed, self.offset = float(p), int(seed), int(offset)

    def __call__(self, site: int, shape, row_stride=None) -> np.ndarray:
        shape = tuple(int(s) for s in shape)
        if row_stride is None:
            return factors(int(np.prod(shape)), self.p, self.seed, self.offset, site).reshape(shape)
        rows = int(np.prod(shape[:-1]))
        f = factors(rows * int(row_stride), self.p, self.seed, self.offset, site).reshape(rows, int(row_stride))
        return np.ascontiguousarray(f[:, : shape[-1]]).reshape(shape)
