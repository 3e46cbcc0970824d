set -x
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -c 3000 gpurun_out/bench.log
