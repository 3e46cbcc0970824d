set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
python scripts/ubench_ops.py --no-rowmax --timeline > gpurun_out/ubench_timeline.log 2>&1; grep "potf2 phases\|potrf n=\|potrs\|laplace_fit max_iter=100\|rff_fit max_iter=100\|gemv" gpurun_out/ubench_timeline.log | head -60
grep -A 42 "potrf timeline n=5000" gpurun_out/ubench_timeline.log | awk 'NR==1||NR%5==2'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocktri|blockrow|blockcol|gemv_kernel" --csv --log-file gpurun_out/solve_launches.csv python scripts/prof_solve.py > gpurun_out/prof_solve.log 2>&1
tail -1 gpurun_out/prof_solve.log
python scripts/ncu_summary.py launches gpurun_out/solve_launches.csv | head -8
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 700 gpurun_out/bench.log
