python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
PPBO_TRACE=1 python scripts/fit_probe.py ackley20d 2> gpurun_out/rfftrace.err | tail -2
grep "rff_fit\]" gpurun_out/rfftrace.err | tail -22 | cut -c1-100
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 300 gpurun_out/bench.log
