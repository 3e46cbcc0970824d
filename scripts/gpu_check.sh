set -x
nvidia-smi -L
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
python scripts/ubench_ops.py > gpurun_out/ubench_ops.log 2>&1; cat gpurun_out/ubench_ops.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 2500 gpurun_out/bench.log
