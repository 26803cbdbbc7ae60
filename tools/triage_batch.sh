cd /root/repo
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -3 gpurun_out/r2v_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
print("rtfx", [r["ms"] for r in d.get("rtfx_per_exit", [])], "parity", d.get("parity_vs_reference", {}).get("ok"), d.get("parity_vs_reference", {}).get("logprob_rel_err_per_exit"))
for k in ("train_with_dropout", "dp_18_layers", "aed_mode"): print(k, d.get(k, {}).get("ms_per_step"))
PY
