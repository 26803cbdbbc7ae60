cd /root/repo
EEC_GEMM_WS=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -4
for w in 2; do for k in 0 2; do EEC_WS_KNOBS=$k EEC_GEMM_WS=$w timeout 120 python tools/gemm_triage.py all 2>&1 | sed "s/^/ws=$w knobs=$k /"; done; done
for w in 2; do for c in silu qkv; do echo "== ws=$w $c"; EEC_GEMM_WS=$w EEC_LIB=early-exit-transformer_b200/eec/libeec_tl.so EEC_GEMM_TL=1 timeout 120 python tools/gemm_triage.py $c 2>&1 | grep -A2 "gemm_ws" | sed -n 17,19p | grep -v "k-block"; done; done
