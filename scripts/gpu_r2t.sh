#!/bin/bash
# last sanity pass of the round: smoke + the kernel / training / model test files
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py tests/test_dropout_gpu.py -m gpu -q -x > gpurun_out/t_tests.log 2>&1
echo "tests exit=$? :: $(tail -n 2 gpurun_out/t_tests.log)"
