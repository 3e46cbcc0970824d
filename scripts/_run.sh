python -m pytest tests/test_full_size.py -m gpu -x -q -k overlapped > gpurun_out/pytest_t11.log 2>&1; tail -15 gpurun_out/pytest_t11.log
