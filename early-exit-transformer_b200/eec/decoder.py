"""CTC decoding beyond greedy: the reference's `--decoder_mode ctc` inference prints the top hypothesis of a CTC
prefix beam search per exit and utterance (util/beam_infer.py:100-110 `BeamInference.ctc_cuda_predict`, called at inference.py:66-79).

``cuda_ctc_decoder`` / ``CUCTCDecoder`` keep the call shape of the torchaudio factory / class the reference uses
(torchaudio/models/decoder/_cuda_ctc_decoder.py): ``decoder(log_prob (B, T, V) on the GPU, encoder_out_lens (B,) int32)`` ->
``List[List[CUCTCHypothesis(tokens, words, score)]]``, `nbest` hypotheses per utterance, best first.  ``decode_all_exits`` is the batched
form the B200 path is built for: all exits and utterances of one forward in ONE kernel launch, one device->host copy at the end.
There is no CPU path: the search runs in libeec.so (csrc/ctc_beam.cu).
"""
from __future__ import annotations

import math
from typing import List, NamedTuple, Sequence, Union

import torch

from . import ops
from .lib import EecError, on_device


class CUCTCHypothesis(NamedTuple):
    tokens: List[int]
    words: List[str]
    score: float


def _vocab(tokens: Union[str, Sequence[str]]) -> List[str]:
    if isinstance(tokens, str):        # a tokens file: first field of every line (torchaudio's _get_vocab_list)
        with open(tokens, "r", encoding="utf-8") as f:
            return [line.strip().split()[0] for line in f]
    return list(tokens)


class CUCTCDecoder:
    def __init__(self, vocab_list: Sequence[str], blank_id: int = 0, beam_size: int = 10, nbest: int = 1,
                 blank_skip_threshold: float = 0.95):
        if blank_id != 0:
            raise AssertionError("blank_id must be 0")
        if not (0 <= blank_skip_threshold <= 1):
            raise AssertionError("blank_skip_threshold must be between 0 and 1")
        self.vocab_list = list(vocab_list)
        self.blank_id, self.nbest = blank_id, nbest
        self.log_skip = math.log(blank_skip_threshold) if blank_skip_threshold > 0 else float("-inf")
        self.beam_size = min(beam_size, len(self.vocab_list))     # (beam size must not exceed the vocabulary, as in torchaudio)
        if self.nbest > self.beam_size:
            raise EecError(f"nbest ({nbest}) must not exceed beam_size ({self.beam_size})")

    def search(self, log_prob: torch.Tensor, encoder_out_lens: torch.Tensor | None = None):
        """Device-side results, no host sync: (tokens [..., nbest, T] int32 padded with -1, n_tokens [..., nbest], scores [..., nbest])
        for log-probabilities of shape (..., T, V)."""
        if not log_prob.is_cuda:
            raise EecError("ctc beam search: log-probabilities must be on a CUDA device (no CPU path)")
        if log_prob.dtype != torch.float32:
            raise EecError("ctc beam search: log-probabilities must be fp32")
        lp = log_prob.contiguous()
        lens = None
        if encoder_out_lens is not None:
            lens = encoder_out_lens.to(device=lp.device, dtype=torch.int32).contiguous()
            if lens.numel() != lp.numel() // (lp.shape[-1] * lp.shape[-2]):
                raise EecError("ctc beam search: one length per emission matrix expected")
        with on_device(lp.device):
            tok, n, sc = ops.ctc_beam_search(lp, lens, self.beam_size, self.nbest, self.blank_id, self.log_skip)
        lead = tuple(lp.shape[:-2])
        return tok.view(*lead, self.nbest, lp.shape[-2]), n.view(*lead, self.nbest), sc.view(*lead, self.nbest)

    def _hyps(self, tok, n, sc) -> List[List[CUCTCHypothesis]]:
        tok, n, sc = tok.cpu(), n.cpu(), sc.cpu()          # the one device->host copy
        out = []
        for b in range(tok.shape[0]):
            row = []
            for j in range(self.nbest):
                ids = tok[b, j, : int(n[b, j])].tolist()
                row.append(CUCTCHypothesis(tokens=ids, words=[self.vocab_list[i] for i in ids], score=float(sc[b, j])))
            out.append(row)
        return out

    def __call__(self, log_prob: torch.Tensor, encoder_out_lens: torch.Tensor) -> List[List[CUCTCHypothesis]]:
        if log_prob.dim() != 3:
            raise EecError("ctc beam search: log_prob must be (batch, frame, num_tokens)")
        return self._hyps(*self.search(log_prob, encoder_out_lens))

    def decode_all_exits(self, out_ebtv: torch.Tensor) -> List[List[List[CUCTCHypothesis]]]:
        """inference.py:66-79 for a whole forward: `out_ebtv` = Early_conformer.forward's (E, B, T', V) log-probs; every exit is decoded
        over ALL T' frames (`enc_len = T'` for every row, util/beam_infer.py:106-107).  One launch; result[e][b] = nbest hypotheses."""
        if out_ebtv.dim() != 4:
            raise EecError("decode_all_exits: expected (E, B, T, V) log-probs")
        tok, n, sc = self.search(out_ebtv, None)
        E = out_ebtv.shape[0]
        return [self._hyps(tok[e], n[e], sc[e]) for e in range(E)]


def cuda_ctc_decoder(tokens: Union[str, Sequence[str]], nbest: int = 1, beam_size: int = 10,
                     blank_skip_threshold: float = 0.95) -> CUCTCDecoder:
    """Same factory signature as torchaudio.models.decoder.cuda_ctc_decoder (util/beam_infer.py:82-83, :108-109)."""
    return CUCTCDecoder(vocab_list=_vocab(tokens), beam_size=beam_size, nbest=nbest, blank_skip_threshold=blank_skip_threshold)


def ctc_cuda_predict(emission: torch.Tensor, tokens: Union[str, Sequence[str]], beam_size: int = 10):
    """BeamInference.ctc_cuda_predict (util/beam_infer.py:100-110): full-length decode of one exit's emissions (B, T', V)."""
    enc_len = torch.full((emission.size(0),), emission.size(1), dtype=torch.int32, device=emission.device)
    return cuda_ctc_decoder(tokens, nbest=1, beam_size=beam_size, blank_skip_threshold=0.95)(emission, enc_len)
