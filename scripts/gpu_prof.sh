set -x
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/potrf_launches.csv python scripts/prof_potrf.py 5000 > gpurun_out/prof_potrf.log 2>&1
tail -2 gpurun_out/prof_potrf.log
