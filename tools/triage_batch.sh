cd /root/repo
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 200 python tools/kbench.py gemm 2>&1 | grep -E "GLU"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-deep --skip-aed > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3d_bench.json").read().strip().splitlines()[-1])
print("step", d["ms_per_step"], "launches", d["gpu_launches"], "rtfx ms", [r["ms"] for r in d["rtfx_per_exit"]], "dropout step", (d.get("train_with_dropout") or {}).get("ms_per_step"), "early exit", (d.get("early_exit_inference") or {}).get("ms"))
PY
