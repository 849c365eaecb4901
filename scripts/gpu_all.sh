#!/bin/bash
# kernel tests + model tests + microbench + bench, logs under gpurun_out/
mkdir -p gpurun_out
bash scripts/gpu_kernel_tests.sh
bash scripts/gpu_full.sh
timeout 600 python scripts/bench_kernels.py --iters 10 > gpurun_out/bench_kernels.log 2>&1
echo "bench_kernels exit=$?"
