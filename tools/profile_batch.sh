#!/bin/bash
# round-2 evidence batch (one GPU): default bench -> launch list of one replayed step -> ncu --set full, one launch per kernel family
cd "$(dirname "$0")/.."
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err || { tail -5 gpurun_out/r2z_bench.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off --csv --log-file gpurun_out/r2z_launches.csv \
  python bench.py --steps 1 --warmup 1 --profile --skip-cpu --skip-rtfx --skip-deep --skip-aed > gpurun_out/r2z_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2z_launches.csv 60 > gpurun_out/r2z_launches_summary.txt 2>&1; head -30 gpurun_out/r2z_launches_summary.txt
KBENCH_PROFILE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2z_kernels \
  python tools/kbench.py attn gemm ln lnbwd conv ctc > gpurun_out/r2z_ncu_kernels.log 2>&1; tail -2 gpurun_out/r2z_ncu_kernels.log
timeout 300 python tools/kbench.py attn gemm ln lnbwd conv ctc beam dec > gpurun_out/r2z_kbench.txt 2>&1; cat gpurun_out/r2z_kbench.txt
