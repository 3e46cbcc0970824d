set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --profile --steps 1 > gpurun_out/prof1.log 2>&1
tail -2 gpurun_out/prof1.log
ncu --set full --clock-control none --import-source on -k regex:ozaki_rowmax -s 1 -c 1 -o gpurun_out/r01d_ozaki python bench.py --profile --steps 1 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
ncu --set full --clock-control none -k regex:"ozaki_slice|ozaki_rowscale|ozaki_merge|blocktri_gemv|blockrow_update|blockcol_update|blockinv_init|sample_omega|potf2_inv|kernel_matrix|gemv_kernel" -s 60 -c 24 -o gpurun_out/r01d_small python bench.py --profile --steps 1 > gpurun_out/prof3.log 2>&1
tail -2 gpurun_out/prof3.log
ls -la gpurun_out/*.ncu-rep
