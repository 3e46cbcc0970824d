set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -5 gpurun_out/pytest_full.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python scripts/ubench_ops.py --no-rowmax --timeline > gpurun_out/ubench_timeline.log 2>&1; grep "potf2 phases\|potrf n=\|laplace_fit max_iter=100\|rff_fit max_iter=100\|total" gpurun_out/ubench_timeline.log | head -20
PPBO_TRACE=1 timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -c 2500 gpurun_out/bench.log
