#!/bin/bash
# usage: gpu_multi.sh N  -> sampling and training benches on N GPUs of one node (torchrun), logs under gpurun_out/
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "sample N=$N exit=$? :: $(tail -n 1 gpurun_out/bench_n$N.log | head -c 400)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 6 --warmup 3 > gpurun_out/bench_train_n$N.log 2> gpurun_out/bench_train_n$N.err
echo "train N=$N exit=$? :: $(tail -n 1 gpurun_out/bench_train_n$N.log | head -c 400)"
tail -n 3 gpurun_out/bench_train_n$N.err
