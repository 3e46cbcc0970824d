python scripts/ubench_ops.py --no-rowmax --timeline --potrf-only > gpurun_out/ubench_potf2.log 2>&1; grep "potrf n=" gpurun_out/ubench_potf2.log | head -20
grep -A 42 "potrf timeline n=5000" gpurun_out/ubench_potf2.log | awk 'NR==1||NR%4==2'
