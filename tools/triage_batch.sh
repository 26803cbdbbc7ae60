cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "layernorm" 2>&1 | tail -2
for b in 296 444 592; do echo "EEC_LNB_BLOCKS=$b"; EEC_LNB_BLOCKS=$b timeout 100 python tools/kbench.py lnbwd 2>&1 | tail -2; done
