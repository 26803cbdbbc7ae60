cd /root/repo
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
