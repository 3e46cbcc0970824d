for R in 20 26; do
PPBO_OVERLAP_RESERVE=$R python bench.py --steps 10 --warmup 3 --no-api-leg --no-cpu-baseline > gpurun_out/bench_t19_R$R.log 2>&1
done
