set -x
python -m pytest tests/test_gpu_ops.py tests/test_src_gpu.py -m gpu -q -x 2>&1 | tail -4
python scripts/ubench_ops.py --no-rowmax 2>&1 | grep -v "  cfg\|rowmax\|gemm_nt\|trsm"
timeout 400 python scripts/fit_overlap_probe.py 2>&1 | head -3
