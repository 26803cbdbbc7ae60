cd /root/repo
for rb in 128 96 64 0; do echo "== EEC_LNP_RB=$rb"; EEC_LNP_RB=$rb timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "layernorm_tail or relu" 2>&1 | tail -1; EEC_LNP_RB=$rb timeout 200 python tools/kbench.py gemm 2>&1 | grep -E "LN"; done
