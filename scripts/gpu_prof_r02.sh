set -x
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -c 300 gpurun_out/bench_n1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_cold.csv python bench.py --profile cold --steps 1 > gpurun_out/prof1.log 2>&1; tail -2 gpurun_out/prof1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_steady.csv python bench.py --profile steady --steps 2 > gpurun_out/prof2.log 2>&1; tail -2 gpurun_out/prof2.log
