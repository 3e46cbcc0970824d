set -x
ncu --set full --clock-control none --import-source on -k regex:ozaki_rowmax -s 1 -c 1 -o gpurun_out/oz_rowmax python scripts/ozaki_probe.py 8192 > gpurun_out/oz_ncu.log 2>&1
tail -3 gpurun_out/oz_ncu.log
ls -la gpurun_out/*.ncu-rep
