#!/bin/bash
# Runs the GPU kernel parity tests group by group, each in its own process with its own timeout, so a
# trapping / hanging kernel in one group cannot poison the CUDA context of the others.
# Usage (on the GPU box): bash scripts/gpu_kernel_tests.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, -k expression
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$2" > "gpurun_out/kt_$1.log" 2>&1
  echo "$1 exit=$? :: $(tail -n 1 gpurun_out/kt_$1.log)" | tee -a gpurun_out/kernel_tests_summary.txt
}
run elementwise "cfg_euler or adaln or patchify or timestep"
for c in cg1n256 cg1n128 cg2n256 cg2n128; do
  run gemm_f32_$c "gemm_bias_f32 and $c"
  run gemm_epi_$c "(gemm_bias_act or gemm_qkv_rope or gemm_gate or gemm_unpatchify) and $c"
done
run gemm_misc "gemm_headline or gemm_rejects"
run attention "gqa_attention"
