set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -5 gpurun_out/pytest_full.log
python scripts/ubench_ops.py --no-rowmax --timeline --potrf-only > gpurun_out/ubench_potf2.log 2>&1; grep "potf2 phases\|potrf n=\|warp potrf32" gpurun_out/ubench_potf2.log | head -60
grep -A 42 "potrf timeline n=5000" gpurun_out/ubench_potf2.log | awk 'NR==1||NR%4==2'
