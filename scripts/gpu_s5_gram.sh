python scripts/gram_probe.py 2>&1 | tail -12
python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "kernel or predict" 2>&1 | tail -3
