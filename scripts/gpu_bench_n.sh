# usage: bash scripts/gpu_bench_n.sh N   (on a box with N GPUs): one bench line into gpurun_out/bench_n$N.log
set -x
n=$1
if [ $n -eq 1 ]; then timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; fi
tail -c 1500 gpurun_out/bench_n$n.log
