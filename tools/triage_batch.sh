cd /root/repo
for w in 1 2; do EEC_GEMM_WS=$w timeout 120 python tools/gemm_triage.py all 2>&1 | sed "s/^/ws=$w /"; done
for w in 1 2; do for c in silu qkv; do echo "== ws=$w $c"; EEC_GEMM_WS=$w EEC_LIB=early-exit-transformer_b200/eec/libeec_tl.so EEC_GEMM_TL=1 timeout 120 python tools/gemm_triage.py $c 2>&1 | grep -A2 "gemm_ws" | sed -n 17,19p | grep -v "k-block"; done; done
echo "== ws=2 no MMA"; EEC_WS_KNOBS=8 EEC_GEMM_WS=2 EEC_LIB=early-exit-transformer_b200/eec/libeec_tl.so EEC_GEMM_TL=1 timeout 120 python tools/gemm_triage.py silu 2>&1 | grep -A2 "gemm_ws" | sed -n 17,19p | grep -v "k-block"
