"""Generate tests/golden/ctc_beam_ref.npz: outputs of the REAL torchaudio `cuda_ctc_decoder` (the library the reference calls at
util/beam_infer.py:100-110) on seeded emissions.  The library is CUDA-only, so this runs on a GPU box:

    gpurun -- 'python oracle/make_beam_golden.py gpurun_out/ctc_beam_ref.npz'      # then copy the file to tests/golden/

The fixture holds the cases' generator arguments (emissions are regenerated from the seed by oracle.ctc_beam_oracle.synthetic_emissions)
and, per utterance, ALL `beam_size` hypotheses the library returns (tokens, lengths, scores).  The script also prints how the CPU
restatement (oracle/ctc_beam_oracle.py) compares, so that a change of library behaviour is visible when the fixture is regenerated.
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ctc_beam_oracle as BO  # noqa: E402

# (seed, B, T, sharp, blank_bias, enc_len mode)
CASES = [
    (1, 6, 40, 4.0, 3.0, "full"),
    (2, 6, 75, 2.0, 2.0, "full"),
    (3, 4, 120, 1.0, 4.0, "full"),       # flat distributions: many near ties
    (4, 4, 60, 6.0, 1.0, "ragged"),      # peaky, varying valid lengths
    (5, 3, 374, 3.0, 3.5, "full"),       # the benchmark's T'
    (6, 2, 30, 4.0, 30.0, "full"),       # (almost) every frame blank-skipped
]
BEAM = 5          # torchaudio 2.11.0's cuda_ctc_decoder faults on B200 (sm_100) for beam_size >= 7 -- the reference's default 10 included
                  # (tools/probe_torchaudio_ctc.py, profiles/r02_torchaudio_ctc_decoder_on_b200.txt); 5 is what the library can pin
V = 256


def library_decode(lp: torch.Tensor, lens: torch.Tensor, beam: int, thr: float, nbest=None):
    from torchaudio.models.decoder import cuda_ctc_decoder
    dec = cuda_ctc_decoder([str(i) for i in range(lp.shape[2])], nbest=nbest or beam, beam_size=beam, blank_skip_threshold=thr)
    res = dec(lp, lens)
    return [[(list(h.tokens), float(h.score)) for h in utt] for utt in res]


def run_case_isolated(k):
    """one case in a child process: a CUDA fault inside the library must not take the other cases down"""
    import pickle
    import subprocess
    import tempfile
    for nbest in (BEAM, 1):
        with tempfile.NamedTemporaryFile(suffix=".pkl") as f:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(k), f.name, str(nbest)], capture_output=True, text=True)
            if r.returncode != 0:
                print(f"case {k} nbest {nbest}: library FAILED rc={r.returncode}: {r.stderr.strip().splitlines()[-1][:300] if r.stderr.strip() else ''}")
                continue
            return pickle.load(open(f.name, "rb"))
    return None


def child(k, path, nbest):
    import pickle
    seed, B, T, sharp, bias, mode = CASES[k]
    lp = BO.synthetic_emissions(B, T, V, seed, sharp, bias)
    lens = case_lens(k)
    lib = library_decode(torch.from_numpy(lp).cuda(), torch.from_numpy(lens).cuda(), BEAM, 0.95, nbest)
    torch.cuda.synchronize()
    pickle.dump(lib, open(path, "wb"))


def case_lens(k):
    seed, B, T, sharp, bias, mode = CASES[k]
    lens = np.full((B,), T, dtype=np.int32)
    if mode == "ragged":
        lens = np.random.RandomState(seed).randint(T // 3, T + 1, size=B).astype(np.int32)
    return lens


def main(path):
    out = {"beam": BEAM, "V": V, "cases": np.array([[c[0], c[1], c[2]] for c in CASES], dtype=np.int64),
           "sharp": np.array([c[3] for c in CASES], dtype=np.float32), "blank_bias": np.array([c[4] for c in CASES], dtype=np.float32),
           "ragged": np.array([c[5] == "ragged" for c in CASES])}
    tot = top1_ok = all_ok = 0
    for k, (seed, B, T, sharp, bias, mode) in enumerate(CASES):
        lp = BO.synthetic_emissions(B, T, V, seed, sharp, bias)
        lens = case_lens(k)
        lib = run_case_isolated(k)
        out[f"ok{k}"] = lib is not None
        if lib is None:
            continue
        tok = np.full((B, BEAM, T), -1, dtype=np.int32)
        ln = np.zeros((B, BEAM), dtype=np.int32)
        sc = np.zeros((B, BEAM), dtype=np.float32)
        for b in range(B):
            for j, (t_, s_) in enumerate(lib[b]):
                tok[b, j, : len(t_)] = t_
                ln[b, j], sc[b, j] = len(t_), s_
        out[f"lens{k}"], out[f"tokens{k}"], out[f"ntok{k}"], out[f"score{k}"] = lens, tok, ln, sc
        mine = BO.decode_batch(lp, lens, BEAM, 0, 0.95, nbest=BEAM)
        for b in range(B):
            tot += 1
            t1 = mine[b][0][0] == lib[b][0][0]
            top1_ok += t1
            same = BO.same_beam(mine[b][: len(lib[b])], lib[b][: len(mine[b])])
            all_ok += same
            if not t1 or not same:
                print(f"case {k} utt {b}: top1 {'ok' if t1 else 'DIFF'} all {'ok' if same else 'DIFF'}")
                for j in range(min(4, len(mine[b]))):
                    print("   lib ", lib[b][j][0][:12], round(lib[b][j][1], 5), "| mine", mine[b][j][0][:12], round(mine[b][j][1], 5))
            ds = max(abs(m[1] - l[1]) for m, l in zip(mine[b], lib[b])) if same else float("nan")
            if same and ds > 1e-3:
                print(f"case {k} utt {b}: score diff {ds}")
    print(f"oracle vs torchaudio cuda_ctc_decoder: top-1 tokens equal {top1_ok}/{tot}, all {BEAM} hypotheses equal {all_ok}/{tot}")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--case":
        child(int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else BEAM)
        sys.exit(0)
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "tests", "golden", "ctc_beam_ref.npz"))
