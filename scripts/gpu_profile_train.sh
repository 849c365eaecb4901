#!/bin/bash
# ncu evidence for the TRAINING step (configs[3]): launch list of every kernel of the run (torch's included) + one
# `--set full` capture of the backward kernels.  Only after the same command exited 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --mode train --steps 1 --warmup 1"
$CMD > gpurun_out/plain_train.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/ncu_train_launches.log 2>&1
echo "launch list exit=$?"
K="regex:adaln_gate_bwd|adaln_bwd_dx_kernel|adaln_bwd_colsum_kernel|gate_bwd_kernel|gqa_attention_bwd_kernel|attn_bwd_rowdot_kernel|attn_bwd_dq_finalize_kernel|colsum_bf16_kernel|adamw_kernel|grad_sumsq_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 30 -c 12 -f -o gpurun_out/prof_train_bwd $CMD > gpurun_out/ncu_train_bwd.log 2>&1
echo "bwd capture exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:adamw_kernel|grad_sumsq_kernel" -c 2 -f -o gpurun_out/prof_train_opt $CMD > gpurun_out/ncu_train_opt.log 2>&1
echo "opt capture exit=$?"
ls -la gpurun_out/ | tail -8
