python -m pytest tests -m gpu -x -q > gpurun_out/pytest_t18.log 2>&1; tail -2 gpurun_out/pytest_t18.log
python bench.py --steps 10 --warmup 3 --no-api-leg --no-cpu-baseline > gpurun_out/bench_t18.log 2>&1
PPBO_CHORD_SPACE=alpha python bench.py --steps 10 --warmup 3 --no-api-leg --no-cpu-baseline > gpurun_out/bench_t18a.log 2>&1
