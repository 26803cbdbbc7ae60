cd /root/repo
timeout 900 python -m pytest tests -q -x -m gpu -k "ctc or golden or headline or step or train" 2>&1 | tail -4
for r in 0 1; do echo "EEC_CTC_RING=$r"; EEC_CTC_RING=$r python tools/ctc_split.py 2>&1 | tail -2; done
