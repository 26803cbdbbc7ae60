cd /root/repo
timeout 600 python -m pytest tests -q -x -m gpu -k "attention or attn or aed or decoder or drop" 2>&1 | tail -2
timeout 200 python tools/kbench.py attn dec 2>&1 | tail -5
EEC_LIB=early-exit-transformer_b200/eec/libeec_tl.so timeout 200 python tools/kbench.py attn 2>&1 | grep -A3 "attn_bwd CTA" | tail -4 | cut -c1-420
