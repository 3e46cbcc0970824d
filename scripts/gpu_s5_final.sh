set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 600 gpurun_out/bench_default.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --profile --steps 1 > gpurun_out/prof1.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r01g_launches.csv > gpurun_out/r01g_launches.txt 2>&1; head -14 gpurun_out/r01g_launches.txt
