#!/bin/bash
# flakiness probe of one test under the two adaln_gate_bwd variants
mkdir -p gpurun_out
for v in 1 0 1 0 1 0 1 1; do
  JAT_AGB_STAGED=$v timeout 300 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k zero_copy_gradient_handoff > gpurun_out/flaky_$v.log 2>&1
  echo "STAGED=$v exit=$? :: $(grep -h 'AssertionError: \|passed\|failed' gpurun_out/flaky_$v.log | head -3 | cut -c1-200)"
done
