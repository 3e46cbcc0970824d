set -x
nvidia-smi -L
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_mgpu.log 2>&1; tail -15 gpurun_out/pytest_mgpu.log
N=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-api-leg > gpurun_out/bench_n$n.log 2>&1
    else timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; fi
    tail -c 3500 gpurun_out/bench_n$n.log
  fi
done
