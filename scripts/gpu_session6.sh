set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_s6.log 2>&1; tail -25 gpurun_out/pytest_s6.log
python scripts/steady_probe.py ackley20d 8 > gpurun_out/steady_ackley.log 2>&1; grep "^append\|warm vs\|GPState\|mu_pred" gpurun_out/steady_ackley.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s6.log 2>&1; tail -c 6500 gpurun_out/bench_s6.log
