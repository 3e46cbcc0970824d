for c in 0.25 0.45; do
PPBO_RFF_CHORD_REL=$c PPBO_TRACE=1 python scripts/steady_probe.py 2>&1 | grep "RFFState.cold\|ppbo_rff_fit\] it" | head -40 | cut -c1-110 > gpurun_out/rffcold_$c.txt
done
