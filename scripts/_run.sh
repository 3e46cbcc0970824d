ncu --set full --clock-control none --import-source on -k regex:kernel_matrix_stream -c 12 -o gpurun_out/r02_gram python scripts/gram_probe.py > gpurun_out/prof_gram.log 2>&1; tail -3 gpurun_out/prof_gram.log
ls -la gpurun_out/r02_gram.ncu-rep
