set -x
PPBO_TRACE=1 python scripts/extend_probe.py ackley20d > gpurun_out/extend_probe_a.log 2>&1; tail -30 gpurun_out/extend_probe_a.log
python scripts/extend_probe.py levy10d > gpurun_out/extend_probe.log 2>&1; tail -12 gpurun_out/extend_probe.log
python scripts/steady_probe.py ackley20d 8 > gpurun_out/steady_ackley.log 2>&1; tail -16 gpurun_out/steady_ackley.log
python -m pytest tests/test_full_size.py tests/test_incremental.py tests/test_gpu_ops.py -m gpu -q > gpurun_out/pytest_s3.log 2>&1; tail -30 gpurun_out/pytest_s3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s3.log 2>&1; tail -c 6000 gpurun_out/bench_s3.log
