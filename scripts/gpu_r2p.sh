#!/bin/bash
# heads per attention CTA (JAT_ATTN_GS) on the training step
mkdir -p gpurun_out
for v in 5 4 3 2 5 4 3; do
  JAT_ATTN_GS=$v timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/train_gs$v.log 2> gpurun_out/train_gs$v.err || tail -3 gpurun_out/train_gs$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/train_gs$v.log').read().strip().splitlines()[-1])
t=d.get('train', d)
print('GS=$v train ms/step', t['ms_per_step'], 'attn_fwd', t['kernels']['gqa_attention_fwd']['ms_per_step'], 'loss', t['loss'], 'clk', t['clocks']['sm_mhz'])
PY
done
