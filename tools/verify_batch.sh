#!/bin/bash
# one-shot verification batch for a GPU box: A/B of the opt-in kernels, the full GPU suite, the bench (stdout = compact report)
cd "$(dirname "$0")/.."
echo "== dwconv A/B (kbench conv)"
for v in 1 2; do echo "EEC_DW_PIPE=$v"; EEC_DW_PIPE=$v timeout 100 python tools/kbench.py conv 2>&1 | head -2; done
echo "== attention A/B (kbench attn)"
for v in 0 1; do echo "EEC_ATTN_FWD8=$v"; EEC_ATTN_FWD8=$v timeout 100 python tools/kbench.py attn 2>&1 | head -2; done
echo "== tests with the opt-in kernels"
EEC_DW_PIPE=2 EEC_ATTN_FWD8=1 timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests, defaults"
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== bench, defaults"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; tail -2 gpurun_out/bench_v.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v.json"))
print("step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
for k in ("train_with_dropout", "early_exit_inference", "fbank_frontend", "torch_eager_same_gpu", "cpu_baseline"):
    print(k, d.get(k))
print("rtfx", [r["ms"] for r in d.get("rtfx_per_exit", [])])
PY
echo "== bench, opt-in kernels"
EEC_DW_PIPE=2 EEC_ATTN_FWD8=1 timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_v2.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v2.json"))
print("step", d["ms_per_step"], "e2e", d["e2e"]["value"], "dropout", d.get("train_with_dropout", {}).get("ms_per_step"), "rtfx", [r["ms"] for r in d.get("rtfx_per_exit", [])])
PY
