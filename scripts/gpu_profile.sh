#!/bin/bash
# ncu evidence for the step: launch list (per-launch device time) + one `--set full` capture of the hot
# kernels.  Run ONLY after the same bench command exited 0 without ncu (done first, below).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --sub none"
K="regex:gemm_tcgen05_kernel|gqa_attention_fwd_kernel|adaln_norm_modulate_kernel|patchify_cast_kernel|cfg_euler_update_kernel|timestep_features_kernel"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
# gemm launches in order: 3 modulation, 2 patch-embed, then per block qkv, out_proj, fc1, fc2
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 5 -c 4 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gqa_attention_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture exit=$?"
ncu --set full --clock-control none --import-source on -k regex:adaln_norm_modulate_kernel -s 2 -c 1 -f -o gpurun_out/prof_adaln $CMD > gpurun_out/ncu_adaln.log 2>&1
echo "adaln capture exit=$?"
ls -la gpurun_out/
