cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -5
for w in 0 1; do EEC_GEMM_WS=$w timeout 120 python tools/gemm_triage.py all 2>&1 | sed "s/^/ws=$w /"; done
