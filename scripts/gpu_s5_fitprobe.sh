for ex in 1 0; do
PPBO_TRACE=1 PPBO_CHORD_EXTRAPOLATE=$ex python scripts/fit_probe.py ackley20d 2> gpurun_out/fit_ex$ex.err | tail -1
tail -16 gpurun_out/fit_ex$ex.err | cut -c1-100
done
for cfg in levy10d hartmann6d camel2d; do for ex in 1 0; do PPBO_CHORD_EXTRAPOLATE=$ex python scripts/fit_probe.py $cfg | tail -1; done; done
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
