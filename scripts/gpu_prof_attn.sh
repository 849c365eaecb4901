#!/bin/bash
# one `ncu --set full` capture of the attention kernel at the headline shape (after the same command passed plain)
mkdir -p gpurun_out
CMD="python scripts/bench_kernels.py --only attn --iters 3"
$CMD > gpurun_out/plain_attn.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_attn.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gqa_attention_fwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_attn3 $CMD > gpurun_out/ncu_attn3.log 2>&1
echo "attn capture exit=$?"; tail -3 gpurun_out/ncu_attn3.log
