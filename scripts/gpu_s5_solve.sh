set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocktri|blockrow|blockcol|gemv_kernel" --csv --log-file gpurun_out/solve_launches.csv python scripts/prof_solve.py > gpurun_out/prof_solve.log 2>&1
tail -2 gpurun_out/prof_solve.log
python scripts/ncu_summary.py gpurun_out/solve_launches.csv | head -12
python scripts/ubench_ops.py --no-rowmax --timeline --potrf-only > gpurun_out/ubench_potf2.log 2>&1; grep "potf2 phases\|potrf n=\|warp potrf32" gpurun_out/ubench_potf2.log | head -60
