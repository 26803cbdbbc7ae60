cd /root/repo
timeout 1500 python -m pytest tests -q -x -m gpu -k "drop or gemm" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-deep --skip-aed --skip-rtfx > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3b_bench.json").read().strip().splitlines()[-1])
print("step", d["ms_per_step"], "dropout step", (d.get("train_with_dropout") or {}))
PY
