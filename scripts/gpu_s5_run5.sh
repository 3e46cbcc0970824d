set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
python scripts/fit_probe.py ackley20d | tail -1
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 900 gpurun_out/bench.log
