set -x
timeout 600 python -m pytest tests/test_ozaki.py -m gpu -x -q 2>&1 | tail -25
timeout 300 python scripts/ozaki_probe.py 2>&1 | tail -12
