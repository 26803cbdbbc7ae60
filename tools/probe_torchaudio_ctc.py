"""Does torchaudio's cuda_ctc_decoder (util/beam_infer.py:100-110's library) run on this GPU at all?  Each configuration in a child process."""
import subprocess
import sys

CHILD = r'''
import sys, torch, numpy as np
V, beam, B, T, thr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
from torchaudio.models.decoder import cuda_ctc_decoder
g = torch.Generator().manual_seed(0)
lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * 3, dim=-1).cuda().contiguous()
lens = torch.full((B,), T, dtype=torch.int32).cuda()
dec = cuda_ctc_decoder([str(i) for i in range(V)], nbest=1, beam_size=beam, blank_skip_threshold=thr)
r = dec(lp, lens)
torch.cuda.synchronize()
print("OK", r[0][0].tokens[:10], r[0][0].score)
'''
import torch, torchaudio
print("torch", torch.__version__, "torchaudio", torchaudio.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
for cfg in [(256, 10, 4, 50, 0.95), (256, 10, 1, 50, 0.95), (500, 10, 4, 50, 0.95), (1024, 10, 4, 50, 0.95), (5000, 10, 2, 50, 0.95),
            (256, 5, 4, 50, 0.95), (256, 10, 4, 50, 1.0), (32, 10, 4, 50, 0.95), (128, 8, 4, 64, 0.95)]:
    r = subprocess.run([sys.executable, "-c", CHILD] + [str(x) for x in cfg], capture_output=True, text=True)
    last = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or [""])[-1][:200]
    print(cfg, "rc", r.returncode, last)
