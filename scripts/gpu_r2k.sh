#!/bin/bash
for cfg in "JAT_X=1" "JAT_GEMM_EW16_DACT=1"; do
  echo "== train $cfg"; env $cfg python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print(' ms/step', d['ms_per_step'], ' '.join(f\"{n}={k[n]['ms_per_step']}\" for n in ('gemm_bias_act','gemm_dact','gemm_accum','attention_bwd')))"
done
JAT_GEMM_EW16_DACT=1 python -m pytest tests/test_dropout_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "dact or dgrad or training_shape" 2>&1 | tail -2
for cfg in "JAT_X=1" "JAT_GEMM_EW16=2"; do
  echo "== sample $cfg"; env $cfg python bench.py --steps 20 --warmup 3 --sub none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print(' ms/step', d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], ' '.join(f\"{n}={k[n]['ms_per_step']}\" for n in ('gemm_bias_act','gemm_gate_residual','gemm_qkv_rope')))"
done
