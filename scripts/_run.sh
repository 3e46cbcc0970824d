python -m pytest tests/test_full_size.py tests/test_incremental.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
python scripts/timeline.py --mode cold --steps 1 --out gpurun_out/timeline_cold.txt --dump gpurun_out/seq_cold.txt 2>&1 | tail -1
