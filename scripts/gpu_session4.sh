set -x
python scripts/ozaki_order_sweep.py > gpurun_out/ozaki_sweep.log 2>&1; cat gpurun_out/ozaki_sweep.log
PPBO_TRACE=1 python scripts/steady_probe.py ackley20d 8 > gpurun_out/steady_ackley.log 2>&1; grep -v "^\[ppbo_rff" gpurun_out/steady_ackley.log | tail -90
python scripts/extend_probe.py levy10d > gpurun_out/extend_probe.log 2>&1; tail -4 gpurun_out/extend_probe.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s4.log 2>&1; tail -c 5000 gpurun_out/bench_s4.log
