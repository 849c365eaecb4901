#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -q -x -k "adaln_gate_bwd or zero_copy or colsum" > gpurun_out/r_tests.log 2>&1
echo "tests exit=$? :: $(tail -n 2 gpurun_out/r_tests.log)"
bash scripts/gpu_profile_train.sh
