cd /root/repo
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
for v in 0 1; do EEC_BF16_DU=$v timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-rtfx --skip-deep --skip-aed > gpurun_out/r2x_bench_$v.json 2> gpurun_out/r2x_bench_$v.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2x_bench_$v.json").read().strip().splitlines()[-1])
print("EEC_BF16_DU=$v step", d["ms_per_step"], "parity", d.get("parity_vs_reference", {}).get("ok"), d.get("parity_vs_reference", {}).get("loss_rel_err"), "dropout step", d.get("train_with_dropout", {}).get("ms_per_step"))
PY
done
