for cr in 0.25 0.5; do
PPBO_TRACE=1 PPBO_CHORD_REL=$cr python scripts/fit_probe.py ackley20d 2> gpurun_out/fit_$cr.err | tail -1
grep -c chord gpurun_out/fit_$cr.err; tail -28 gpurun_out/fit_$cr.err | cut -c1-95
done
for cfg in levy10d hartmann6d camel2d; do for cr in 0.25 0.5; do PPBO_CHORD_REL=$cr python scripts/fit_probe.py $cfg | tail -1; done; done
