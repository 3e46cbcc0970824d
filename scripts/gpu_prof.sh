set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --profile --steps 1 > gpurun_out/prof1.log 2>&1
tail -2 gpurun_out/prof1.log
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_rowmax -c 1 -o gpurun_out/r01_rowmax python bench.py --profile --steps 0 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_store|potf2_inv|trsv_fwd_chain|trsv_bwd_chain|kernel_matrix_kernel|gemv_kernel|newton_matrix|rff_features|sample_omega" -s 40 -c 24 -o gpurun_out/r01_fit python scripts/prof_potrf.py 5000 --fit > gpurun_out/prof3.log 2>&1
tail -2 gpurun_out/prof3.log
ls -la gpurun_out/*.ncu-rep
