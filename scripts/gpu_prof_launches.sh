set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --profile --steps 1 > gpurun_out/prof1.log 2>&1
tail -2 gpurun_out/prof1.log
python scripts/ncu_summary.py gpurun_out/r01b_launches.csv > gpurun_out/r01b_launches.txt 2>&1; head -50 gpurun_out/r01b_launches.txt
