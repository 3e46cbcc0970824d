set -x
python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2
python scripts/ubench_ops.py --no-rowmax --timeline --potrf-only > gpurun_out/ubench_potf2.log 2>&1; grep "potrf n=" gpurun_out/ubench_potf2.log | head
grep -A 42 "potrf timeline n=5000" gpurun_out/ubench_potf2.log | head -24
python scripts/fit_probe.py ackley20d | tail -1
