python scripts/timeline.py --mode gpfit --dump gpurun_out/seq_gpfit.txt 2>&1 | tail -1
python -m pytest tests/test_full_size.py tests/test_gpu_ops.py tests/test_incremental.py tests/test_src_full_size.py -m gpu -x -q > gpurun_out/pytest_t10.log 2>&1; tail -3 gpurun_out/pytest_t10.log
