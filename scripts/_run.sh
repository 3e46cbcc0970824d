timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -c 200 gpurun_out/bench_n1.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
