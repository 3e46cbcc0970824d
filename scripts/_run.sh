python -m pytest tests -m gpu -x -q > gpurun_out/pytest_t15.log 2>&1; tail -3 gpurun_out/pytest_t15.log
