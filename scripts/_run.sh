for i in 1 2 3; do python bench.py --steps 10 --warmup 3 --no-api-leg --no-cpu-baseline > gpurun_out/bench_t24_$i.log 2>&1; done
