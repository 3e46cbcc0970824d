set -x
python scripts/ubench_ops.py --no-rowmax --timeline --potrf-only > gpurun_out/ubench_potf2.log 2>&1; grep "potf2 phases\|potrf n=\|warp potrf32" gpurun_out/ubench_potf2.log | head -60
