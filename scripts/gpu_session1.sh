set -x
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_s1.log 2>&1; tail -40 gpurun_out/pytest_s1.log
python scripts/steady_probe.py ackley20d 6 > gpurun_out/steady_ackley.log 2>&1; tail -30 gpurun_out/steady_ackley.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s1.log 2>&1; tail -c 2500 gpurun_out/bench_s1.log
