"""CPU oracle of the CTC prefix beam search the reference prints with `inference.py --decoder_mode ctc`.  TEST INFRASTRUCTURE ONLY
(imported by tests/ and tools that generate fixtures; never by the product path).

Reference call sites: util/beam_infer.py:100-110 (`ctc_cuda_predict`: `cuda_ctc_decoder(tokens, nbest=1, beam_size=args.beam_size,
blank_skip_threshold=0.95)` applied to one exit's emissions (B, T', V) with `enc_len = T'` for EVERY row -- padded frames are decoded
too), called once per exit at inference.py:66-79, which prints `best_[0].tokens` of every utterance.

The arithmetic lives in a third-party dependency that is absent from /root/reference: torchaudio's `torchaudio_prefixctc` extension
(torchaudio 2.11.0 in this image, un-pinned by the reference; Python wrapper torchaudio/models/decoder/_cuda_ctc_decoder.py), a
compiled CUDA library without sources here.  This file restates its PUBLISHED algorithm -- CTC prefix beam search (Hannun et al. 2014,
"First-pass large vocabulary continuous speech recognition using bi-directional recurrent DNNs", Alg. 1, without language model) with
the wrapper's documented blank-frame skipping -- in plain Python / numpy:

  * frames t < enc_len with log_prob[t, blank] > log(blank_skip_threshold) are "skipped" (wrapper docstring, :70-72 / :139-141): the
    library does not expand or prune on them, but it does account for them as a pure blank emission -- every hypothesis moves its whole
    mass to p_blank times that frame's blank probability (p_b' = (p_b + p_nb) * p(blank), p_nb' = 0).  This is what the library's outputs
    pin (tests/golden/ctc_beam_ref.npz: scores fall by the skipped frames' blank log-probabilities and a repeat after a skipped frame
    starts a new token); dropping such frames outright does NOT reproduce them;
  * every hypothesis (prefix) carries log p_blank and log p_non_blank; a frame extends every prefix by every token, merges equal
    prefixes with log-add, keeps the `beam_size` best by log(p_b + p_nb);
  * the search starts from the empty prefix with probability 1, so the first expanded frame seeds the beam with its `beam_size` most
    probable tokens (blank = the empty prefix);
  * hypotheses are returned best first, score = log(p_b + p_nb).

Parity pin: tests/golden/ctc_beam_ref.npz holds outputs of the REAL `torchaudio.models.decoder.cuda_ctc_decoder` run on a B200
(oracle/make_beam_golden.py, seeded emissions, all `beam_size` hypotheses with scores); tests/test_oracle_golden.py checks this
restatement against them token for token, and the GPU tests compare the CUDA kernel with the library live on the same emissions.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

NEG_INF = float("-inf")


def _lse(a: float, b: float) -> float:
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    m = a if a > b else b
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def prefix_beam_search(lp: np.ndarray, enc_len: int, beam_size: int = 10, blank: int = 0,
                       blank_skip_threshold: float = 0.95, nbest: int = 1) -> List[Tuple[List[int], float]]:
    """lp (T, V) log-probabilities of ONE utterance -> [(tokens, score)] best first (at most nbest).
    float32 inputs are scored in float32 steps like the library (log-add rounded to fp32 after every step)."""
    T, V = lp.shape
    f32 = np.float32
    thr = f32(math.log(blank_skip_threshold)) if blank_skip_threshold > 0 else NEG_INF
    beam_size = min(beam_size, V)
    beams = [((), 0.0, NEG_INF)]          # list of (prefix tuple, pb, pnb): the empty prefix with probability 1
    for t in range(min(int(enc_len), T)):
        row = lp[t]
        if row[blank] > thr:              # blank-dominated frame: pure blank transition for every hypothesis, no expansion
            beams = [(p, float(f32(_lse(pb, pnb) + float(row[blank]))), NEG_INF) for p, pb, pnb in beams]
            continue
        cand = {}                                                       # prefix -> [pb, pnb]

        def add(prefix, pb, pnb):
            e = cand.get(prefix)
            if e is None:
                cand[prefix] = [pb, pnb]
            else:
                e[0], e[1] = _lse(e[0], pb), _lse(e[1], pnb)
        for prefix, pb, pnb in beams:
            tot = _lse(pb, pnb)
            add(prefix, float(f32(tot + float(row[blank]))), NEG_INF)                       # stay by emitting blank
            if prefix:
                add(prefix, NEG_INF, float(f32(pnb + float(row[prefix[-1]]))))              # stay by repeating the last token
            for c in range(V):
                if c == blank:
                    continue
                base = pb if (prefix and c == prefix[-1]) else tot                       # a repeat extends only through a blank
                if base == NEG_INF:
                    continue
                add(prefix + (c,), NEG_INF, float(f32(base + float(row[c]))))
        scored = [(float(f32(_lse(pb, pnb))), p, pb, pnb) for p, (pb, pnb) in cand.items()]
        scored.sort(key=lambda s: -s[0])
        beams = [(p, pb, pnb) for _, p, pb, pnb in scored[:beam_size]]
    out = [(list(p), float(f32(_lse(pb, pnb)))) for p, pb, pnb in beams]
    out.sort(key=lambda s: -s[1])
    return out[:nbest]


def decode_batch(lp_btv: np.ndarray, enc_lens, beam_size: int = 10, blank: int = 0, blank_skip_threshold: float = 0.95, nbest: int = 1):
    """util/beam_infer.py:100-110 for a batch: one list of (tokens, score) per utterance."""
    return [prefix_beam_search(np.asarray(lp_btv[b]), int(enc_lens[b]), beam_size, blank, blank_skip_threshold, nbest)
            for b in range(lp_btv.shape[0])]


def same_beam(a, b, eps: float = 2e-5) -> bool:
    """Tie-aware comparison of two n-best lists [(tokens, score)]: hypotheses whose scores differ by less than `eps` (relative) are an
    unordered cluster -- fp32 summation order decides their rank -- and a cluster cut off by the end of the list only has to be consistent."""
    if len(a) != len(b):
        return False
    i, n = 0, len(a)
    while i < n:
        j = i + 1
        while j < n and abs(b[j][1] - b[i][1]) <= eps * max(1.0, abs(b[i][1])):
            j += 1
        sa, sb = {tuple(h[0]) for h in a[i:j]}, {tuple(h[0]) for h in b[i:j]}
        if sa != sb and j < n:
            return False
        if sa != sb and j == n:          # the last cluster may continue past the cut: scores must still agree
            if any(abs(x[1] - b[i][1]) > eps * max(1.0, abs(b[i][1])) for x in a[i:j]):
                return False
        if any(abs(x[1] - y[1]) > 10 * eps * max(1.0, abs(y[1])) for x, y in zip(a[i:j], b[i:j])):
            return False
        i = j
    return True


def synthetic_emissions(B: int, T: int, V: int = 256, seed: int = 0, sharp: float = 4.0, blank_bias: float = 3.0) -> np.ndarray:
    """Seeded (B, T, V) fp32 log-softmax emissions: Gaussian logits times `sharp` with a blank bias on a random half of the frames, so
    that some frames pass the 0.95 blank-skip threshold, repeats occur and the beam is neither trivial nor flat."""
    rng = np.random.RandomState(seed)
    z = rng.standard_normal((B, T, V)).astype(np.float32) * np.float32(sharp)
    boost = (rng.uniform(size=(B, T)) < 0.5).astype(np.float32) * np.float32(blank_bias * sharp)
    z[:, :, 0] += boost
    hold = rng.uniform(size=(B, T)) < 0.3                              # (nearly) repeat the previous frame's logits: consecutive repeats
    jitter = rng.standard_normal((B, T, V)).astype(np.float32) * np.float32(0.05 * sharp)   # (exact copies make exactly tied hypotheses)
    for t in range(1, T):
        z[:, t][hold[:, t]] = z[:, t - 1][hold[:, t]] + jitter[:, t][hold[:, t]]
    z = z - z.max(axis=2, keepdims=True)
    lse = np.log(np.exp(z.astype(np.float64)).sum(axis=2, keepdims=True))
    return (z.astype(np.float64) - lse).astype(np.float32)
