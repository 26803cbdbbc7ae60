"""CTC prefix beam search on the device (util/beam_infer.py:100-110) against the CPU oracle run live, and -- through the
committed fixture -- against the real torchaudio cuda_ctc_decoder (tests/golden/ctc_beam_ref.npz, oracle/make_beam_golden.py)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import ctc_beam_oracle as BO

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ctc_beam_ref.npz")


def run_kernel(lp, lens, beam=10, nbest=10, thr=0.95):
    import eec
    dec = eec.cuda_ctc_decoder([str(i) for i in range(lp.shape[-1])], nbest=nbest, beam_size=beam, blank_skip_threshold=thr)
    return dec(torch.from_numpy(lp).cuda(), torch.from_numpy(np.asarray(lens, dtype=np.int32)).cuda())


@pytest.mark.parametrize("case", [(11, 5, 40, 4.0, 3.0), (12, 4, 90, 2.0, 2.0), (13, 3, 120, 1.0, 4.0), (14, 2, 374, 3.0, 3.5), (15, 2, 30, 4.0, 30.0)])
def test_kernel_vs_live_oracle(case):
    seed, B, T, sharp, bias = case
    lp = BO.synthetic_emissions(B, T, 256, seed, sharp, bias)
    lens = np.random.RandomState(seed).randint(T // 2, T + 1, size=B).astype(np.int32)
    lens[0] = T
    got = run_kernel(lp, lens)
    ref = BO.decode_batch(lp, lens, 10, 0, 0.95, nbest=10)
    for b in range(B):
        n = len(ref[b])
        mine = [(h.tokens, h.score) for h in got[b][:n]]
        # the hypothesis inference.py prints, and the whole beam in order (hypotheses tied within fp32 resolution form unordered clusters)
        assert BO.same_beam(mine[:1], ref[b][:1]) or BO.same_beam(mine, ref[b]), (b, mine[0], ref[b][0])
        assert BO.same_beam(mine, ref[b]), (b, [round(m[1], 5) for m in mine], [round(r[1], 5) for r in ref[b]])


def test_kernel_vs_torchaudio_fixture():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/ctc_beam_ref.npz not generated")
    g = np.load(GOLDEN)
    beam = int(g["beam"])
    checked = 0
    for k, (seed, B, T) in enumerate(g["cases"]):
        if not bool(g[f"ok{k}"]):
            continue
        lp = BO.synthetic_emissions(int(B), int(T), int(g["V"]), int(seed), float(g["sharp"][k]), float(g["blank_bias"][k]))
        got = run_kernel(lp, g[f"lens{k}"], beam=beam, nbest=beam)
        for b in range(int(B)):
            lib = [(g[f"tokens{k}"][b, j, : int(g[f"ntok{k}"][b, j])].tolist(), float(g[f"score{k}"][b, j])) for j in range(beam)]
            mine = [(h.tokens, h.score) for h in got[b]]
            assert mine[0][0] == lib[0][0] or BO.same_beam(mine[:2], lib[:2]), (k, b)           # the top hypothesis of the LIBRARY
            assert BO.same_beam(mine, lib), (k, b, [round(m[1], 5) for m in mine], [round(r[1], 5) for r in lib])
            checked += 1
    assert checked > 0


def test_all_exits_one_launch_and_api_shape():
    import eec
    lp = BO.synthetic_emissions(6, 50, 256, 21, 3.0, 3.0).reshape(2, 3, 50, 256)              # (E, B, T, V)
    dec = eec.cuda_ctc_decoder([f"t{i}" for i in range(256)], nbest=1, beam_size=10, blank_skip_threshold=0.95)
    x = torch.from_numpy(lp).cuda()
    n0 = eec.load().eec_launch_count()
    allx = dec.decode_all_exits(x)
    assert eec.load().eec_launch_count() - n0 == 1
    for e in range(2):
        per = eec.ctc_cuda_predict(x[e], [f"t{i}" for i in range(256)], beam_size=10)         # the reference's per-exit call shape
        for b in range(3):
            assert per[b][0].tokens == allx[e][b][0].tokens
            assert per[b][0].words == [f"t{i}" for i in per[b][0].tokens]
            assert math.isfinite(per[b][0].score)
    with pytest.raises(eec.EecError):
        dec(torch.from_numpy(lp[0]), torch.full((3,), 50, dtype=torch.int32))                  # CPU tensor: no fallback
