set -x
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -c 600 gpurun_out/bench_n1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_cold.csv python bench.py --profile cold --steps 1 > gpurun_out/prof1.log 2>&1; tail -2 gpurun_out/prof1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_steady.csv python bench.py --profile steady --steps 2 > gpurun_out/prof2.log 2>&1; tail -2 gpurun_out/prof2.log
ncu --set full --clock-control none --import-source on -k regex:ozaki_rowmax -s 1 -c 1 -o gpurun_out/r02_ozaki python bench.py --profile cold --steps 1 > gpurun_out/prof3.log 2>&1; tail -2 gpurun_out/prof3.log
ncu --set full --clock-control none --import-source on -k regex:"skinny_nt|border_solve|border_update|blocktri_gemv|blockrow_update|blockcol_update|gemv_kernel|chord_decide|rff_grad_kernel|rff_anderson|rff_ls_scalars" -s 80 -c 24 -o gpurun_out/r02_steady_fit python bench.py --profile steady --steps 1 > gpurun_out/prof4.log 2>&1; tail -2 gpurun_out/prof4.log
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
