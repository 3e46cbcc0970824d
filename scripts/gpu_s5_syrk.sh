ncu --set full --clock-control none --import-source on -k regex:"syrk128" -s 3 -c 1 -o gpurun_out/r01f_syrk2 python scripts/prof_potrf.py 5000 > gpurun_out/prof4.log 2>&1
tail -1 gpurun_out/prof4.log
