set -x
timeout 600 python -m pytest tests/test_ozaki.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/ozaki_probe.py 2>&1 | tail -12
timeout 300 python scripts/ozaki_probe.py --diag 2>&1 | tail -6
