import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "early-exit-transformer_b200")]
from oracle import ctc_beam_oracle as BO
import eec
for case in [(12, 4, 90, 2.0, 2.0), (14, 2, 374, 3.0, 3.5)]:
    seed, B, T, sharp, bias = case
    lp = BO.synthetic_emissions(B, T, 256, seed, sharp, bias)
    lens = np.random.RandomState(seed).randint(T // 2, T + 1, size=B).astype(np.int32); lens[0] = T
    dec = eec.cuda_ctc_decoder([str(i) for i in range(256)], nbest=10, beam_size=10, blank_skip_threshold=0.95)
    got = dec(torch.from_numpy(lp).cuda(), torch.from_numpy(lens).cuda())
    ref = BO.decode_batch(lp, lens, 10, 0, 0.95, nbest=10)
    for b in range(B):
        print("case", case, "utt", b)
        for j in range(10):
            g, r = got[b][j], ref[b][j]
            same = g.tokens == r[0]
            inref = [k for k in range(10) if ref[b][k][0] == g.tokens]
            print(f"  {j}: kernel {g.score:.6f} len {len(g.tokens)} | oracle {r[1]:.6f} len {len(r[0])} {'same' if same else 'DIFF (kernel hyp is oracle #' + str(inref) + ')'}")
# which beam sizes does the library survive?
import subprocess
CHILD = r'''
import sys, torch
beam = int(sys.argv[1])
from torchaudio.models.decoder import cuda_ctc_decoder
g = torch.Generator().manual_seed(0)
lp = torch.log_softmax(torch.randn(4, 50, 256, generator=g) * 3, dim=-1).cuda().contiguous()
lens = torch.full((4,), 50, dtype=torch.int32).cuda()
r = cuda_ctc_decoder([str(i) for i in range(256)], nbest=1, beam_size=beam, blank_skip_threshold=0.95)(lp, lens)
torch.cuda.synchronize(); print("OK")
'''
ok = []
for beam in range(1, 21):
    r = subprocess.run([sys.executable, "-c", CHILD, str(beam)], capture_output=True, text=True)
    ok.append((beam, r.returncode == 0))
print("torchaudio cuda_ctc_decoder on this GPU, beam -> runs:", ok)
