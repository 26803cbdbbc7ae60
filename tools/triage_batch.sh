cd /root/repo
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
