set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
python scripts/ubench_ops.py --no-rowmax --timeline --potrf-only > gpurun_out/ubench_potf2.log 2>&1; grep "potrf n=" gpurun_out/ubench_potf2.log | head
grep -A 42 "potrf timeline n=5000" gpurun_out/ubench_potf2.log | awk 'NR==1||NR%4==2'
python scripts/fit_probe.py ackley20d | tail -1
