cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r2y_dpcheck.log 2>&1; tail -12 gpurun_out/r2y_dpcheck.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu > gpurun_out/r2y_bench2.json 2> gpurun_out/r2y_bench2.err; tail -2 gpurun_out/r2y_bench2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_bench2.json").read().strip().splitlines()[-1])
print("2 GPUs: step", d["ms_per_step"], "value", d["value"], "dp18", (d.get("dp_18_layers") or {}).get("ms_per_step"))
PY
