#!/bin/bash
# scratch batch for a GPU box (`gpurun -- 'bash tools/triage_batch.sh'`): the full GPU suite and a short bench at HEAD.
# Rewritten freely during kernel work (A/B runs of env knobs, timelines, ncu captures); tools/profile_batch.sh is the recorded evidence batch.
# (Do not chain more than two torchrun jobs in one gpurun call on this pool: the third one hung twice in round 2.)
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-deep --skip-aed > gpurun_out/head_bench.json 2> gpurun_out/head_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/head_bench.json").read().strip().splitlines()[-1])
print("step", d["ms_per_step"], "six-exit forward ms", [r["ms"] for r in d["rtfx_per_exit"]][-1], "dropout step", (d.get("train_with_dropout") or {}).get("ms_per_step"))
PY
