set -x
PPBO_TRACE=1 python scripts/anderson_probe.py ackley20d > gpurun_out/anderson_a.log 2>&1; grep -v "rff_fit" gpurun_out/anderson_a.log | grep "anderson=\|append\|warm vs\|it [0-9]* chord\|newton" | tail -120
python scripts/anderson_probe.py levy10d > gpurun_out/anderson_l.log 2>&1; grep "anderson=\|append\|warm vs" gpurun_out/anderson_l.log
