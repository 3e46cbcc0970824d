set -x
python scripts/extend_probe.py levy10d > gpurun_out/extend_probe.log 2>&1; tail -20 gpurun_out/extend_probe.log
PPBO_TRACE=1 python scripts/extend_probe.py ackley20d > gpurun_out/extend_probe_a.log 2>&1; tail -40 gpurun_out/extend_probe_a.log
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_s2.log 2>&1; tail -60 gpurun_out/pytest_s2.log
