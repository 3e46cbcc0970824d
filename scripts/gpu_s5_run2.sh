set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -5 gpurun_out/pytest_full.log
python scripts/ubench_ops.py --no-rowmax --timeline > gpurun_out/ubench_timeline.log 2>&1; grep "potf2 phases\|potrf n=\|potrs\|laplace_fit max_iter=100\|rff_fit max_iter=100\|total\|gemv\|x256\|x128" gpurun_out/ubench_timeline.log | head -60
PPBO_TRACE=1 timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 1800 gpurun_out/bench.log
