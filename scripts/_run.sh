python scripts/_dbg.py 2>&1 | grep -A1 "levy\|rq3d" | head
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_t8.log 2>&1; tail -3 gpurun_out/pytest_t8.log
for R in 12 20; do
PPBO_OVERLAP_RESERVE=$R python bench.py --steps 8 --warmup 3 --no-api-leg --no-cpu-baseline > gpurun_out/bench_t8_R$R.log 2>&1
done
