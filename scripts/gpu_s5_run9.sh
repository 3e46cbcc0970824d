set -x
python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/pytest_full.log; tail -3 gpurun_out/pytest_full.log
ncu --set full --clock-control none -k regex:"syrk128" -s 3 -c 2 -o gpurun_out/r01f_syrk python scripts/prof_potrf.py 5000 > gpurun_out/prof4.log 2>&1
tail -1 gpurun_out/prof4.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -c 300 gpurun_out/bench.log
