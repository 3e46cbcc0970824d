set -x
python -m pytest tests -m gpu -q 2>&1 > gpurun_out/pytest_full.log; tail -8 gpurun_out/pytest_full.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python scripts/ubench_ops.py > gpurun_out/ubench_ops.log 2>&1; grep -v "  cfg" gpurun_out/ubench_ops.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 1500 gpurun_out/bench.log
