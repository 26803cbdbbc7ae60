"""Length-sorted sub-batching of the reference's collate function and the graph-shape policy that goes with it (second half).

``length_sorted_split`` is the chunking of ``CollatePaddingFn.__call__`` (util/data_loader.py:163-188): the mini-batch is sorted by feature
length, longest first, and cut into ``n_batch_split`` consecutive chunks of roughly equal TOTAL length (so a chunk of long utterances holds
fewer of them); ``train()`` runs one model call + optimiser step per chunk and skips mini-batches that did not produce exactly
``n_batch_split`` chunks (train.py:22-26).  Each chunk is padded to ITS longest utterance only, which is what keeps the padding ratio low --
and what makes the set of ``(B_chunk, T_in)`` shapes open-ended.

``GraphedStepCache`` maps those shapes onto captured CUDA graphs.  Neither the batch size nor the time axis of a chunk can be padded without
changing the arithmetic: train-mode BatchNorm statistics and the loss mean run over all (b, t) of the call, and the reference's own
precondition ``max(lengths) // 4 >= T'`` forbids a time axis longer than the longest utterance.  So a graph is keyed by the exact
``(B_chunk, T_in)``; only the target width (pure padding: ``target_lengths`` governs the CTC recursion) is bucketed.  Graphs are captured on
first use and kept in an LRU of bounded size; a shape that is not (yet) captured runs eagerly through the same kernels.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, List, Sequence


def length_sorted_split(lengths: Sequence[int], n_split: int) -> List[List[int]]:
    """Indices (into the given mini-batch) of every chunk, in the reference's order (util/data_loader.py:164-188)."""
    order = sorted(range(len(lengths)), key=lambda i: lengths[i], reverse=True)     # :164 (stable, like Python's sorted there)
    s_sum = sum(lengths) / n_split                                                  # :167
    chunks, p_sum, init, p_split = [], 0, 0, 0
    for end, i in enumerate(order):                                                 # :174-184
        p_sum += lengths[i]
        if p_sum >= s_sum:
            chunks.append(order[init:end + 1])
            p_sum = 0
            p_split += 1
            init = end + 1
    if p_split != n_split:                                                          # :186-187
        chunks.append(order[init:len(order)])
    return chunks


def trains_on(chunks: List[List[int]], n_split: int) -> bool:
    """train.py:22-24: a mini-batch whose collate produced a different number of chunks is skipped entirely."""
    return len(chunks) == n_split


def padding_ratio(lengths: Sequence[int], chunks: List[List[int]]) -> float:
    """padded frames / real frames when every chunk is padded to its own longest utterance (pad_sequence, util/data_loader.py:223)"""
    real = sum(lengths[i] for c in chunks for i in c)
    padded = sum(max(lengths[i] for i in c) * len(c) for c in chunks if c)
    return padded / max(real, 1) - 1.0


def target_bucket(width: int, step: int = 16) -> int:
    return ((max(int(width), 1) + step - 1) // step) * step


class GraphedStepCache:
    """(B_chunk, T_in, bucketed target width) -> captured training step, LRU-bounded.

        cache = eec.batching.GraphedStepCache(lambda B, T_in, L: eec.GraphedTrainStep(model, B, T_in, L, optimizer=opt), capacity=32)
        for src, targets, t_len, lengths in chunked_batch:            # the reference's c_batch loop, train.py:26-32
            loss = cache.step(src, lengths, targets, t_len)
    """

    def __init__(self, make_step: Callable[[int, int, int], object], capacity: int = 32, target_step: int = 16):
        self.make_step, self.capacity, self.target_step = make_step, int(capacity), int(target_step)
        self.steps: "OrderedDict[tuple, object]" = OrderedDict()
        self.hits = self.captures = self.evictions = 0

    def key(self, batch: int, t_in: int, target_width: int) -> tuple:
        return (int(batch), int(t_in), target_bucket(target_width, self.target_step))

    def get(self, batch: int, t_in: int, target_width: int):
        k = self.key(batch, t_in, target_width)
        st = self.steps.get(k)
        if st is not None:
            self.steps.move_to_end(k)
            self.hits += 1
            return st
        if len(self.steps) >= self.capacity:
            self.steps.popitem(last=False)      # least recently used graph (and its static buffers) is released
            self.evictions += 1
        st = self.make_step(*k)
        self.steps[k] = st
        self.captures += 1
        return st

    def step(self, src, lengths, targets, target_lengths):
        st = self.get(src.shape[0], src.shape[2], targets.shape[1])
        return st(src, lengths, targets, target_lengths)
