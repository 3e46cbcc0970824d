python -m pytest tests/test_full_size.py tests/test_incremental.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
python scripts/timeline.py --mode steady --steps 4 --api --out gpurun_out/timeline_steady.txt --dump gpurun_out/seq_steady.txt 2>&1 | tail -1
python scripts/steady_probe.py 2>&1 | grep "append" | cut -c1-120
