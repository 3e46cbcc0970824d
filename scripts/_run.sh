python scripts/gram_probe.py 2>&1 | tail -12
