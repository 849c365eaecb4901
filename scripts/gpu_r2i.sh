#!/bin/bash
mkdir -p gpurun_out
for cfg in "JAT_X=0" "JAT_PDL=1" "JAT_GEMM_TAIL=2" "JAT_PDL=1 JAT_GEMM_TAIL=2"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --sub none --no-cpu-baseline --no-e2e > gpurun_out/r2i_$tag.log 2> gpurun_out/r2i_$tag.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_$tag.log').read().strip().splitlines()[-1]); k=d['kernels']
print('$tag: ms/step', d['ms_per_step'], 'ksum', d['kernel_sum_ms_per_step'], 'share', d['kernel_time_share_of_step'], 'clk', d['clocks']['sm_mhz'], ' '.join(f"{n[:14]}={k[n]['ms_per_step']:.3f}" for n in ('gemm_gate_residual','gemm_bias_act','gqa_attention_fwd','gemm_qkv_rope','adaln_norm_modulate')))
PY
done
