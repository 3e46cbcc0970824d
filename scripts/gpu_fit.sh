set -x
python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -3
python scripts/ubench_ops.py --no-rowmax --timeline 2>&1 | grep "potf2 phases\|potrf n=\|laplace_fit max_iter=100\|rff_fit max_iter=100\|total"
