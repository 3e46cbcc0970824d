set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 400 gpurun_out/bench.log
