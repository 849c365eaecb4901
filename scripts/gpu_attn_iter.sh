#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gqa_attention" > gpurun_out/kt_attention.log 2>&1
echo "attention tests exit=$? :: $(tail -n 1 gpurun_out/kt_attention.log)"
grep -E "^E  " gpurun_out/kt_attention.log | head -5
timeout 300 python scripts/bench_kernels.py --only attn --iters 10 > gpurun_out/bench_attn.log 2>&1
timeout 120 python scripts/att_trace.py >> gpurun_out/bench_attn.log 2>&1
cat gpurun_out/bench_attn.log
