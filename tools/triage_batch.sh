cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm or wgrad" 2>&1 | tail -4
for w in 0 1; do echo "== EEC_GEMM_PAIR=$w"; EEC_GEMM_PAIR=$w timeout 200 python tools/kbench.py gemm 2>&1 | grep -E "wgrad|dgrad"; done
