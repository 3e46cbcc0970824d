set -x
python scripts/ozaki_order_sweep.py > gpurun_out/ozaki_sweep.log 2>&1; cat gpurun_out/ozaki_sweep.log
PPBO_TRACE=1 python scripts/steady_probe.py ackley20d 8 > gpurun_out/steady_ackley.log 2>&1; grep "^append\|warm vs\|newton" gpurun_out/steady_ackley.log | tail -30
python scripts/steady_probe.py levy10d 6 > gpurun_out/steady_levy.log 2>&1; grep "^append\|warm vs\|cold" gpurun_out/steady_levy.log | tail -12
python -m pytest tests -m gpu -q > gpurun_out/pytest_s5.log 2>&1; tail -15 gpurun_out/pytest_s5.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s5.log 2>&1; tail -c 5500 gpurun_out/bench_s5.log
