ncu --set full --clock-control none -k regex:"kernel_matrix_mma|km_rowsum" -c 6 -o gpurun_out/r01g_gram python scripts/gram_probe.py > gpurun_out/prof5.log 2>&1
tail -2 gpurun_out/prof5.log
