set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --profile --steps 1 > gpurun_out/prof1.log 2>&1
tail -2 gpurun_out/prof1.log
python scripts/ncu_summary.py launches gpurun_out/r01f_launches.csv > gpurun_out/r01f_launches.txt 2>&1; head -12 gpurun_out/r01f_launches.txt
ncu --set full --clock-control none -k regex:"potf2_inv|rowpanel128|blocktri_gemv|blockcol_update|blockrow_update|chord_decide|gemv_kernel" -s 60 -c 14 -o gpurun_out/r01f_fit python scripts/prof_potrf.py 5000 --fit > gpurun_out/prof3.log 2>&1
tail -2 gpurun_out/prof3.log
ls -la gpurun_out/*.ncu-rep
