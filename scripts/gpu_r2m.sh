#!/bin/bash
# staged (bulk-copy) adaln_gate_bwd: kernel tests, micro-benchmark A/B, training-step A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "adaln_gate_bwd" > gpurun_out/agb_tests.log 2>&1
echo "agb_tests exit=$? :: $(tail -n 2 gpurun_out/agb_tests.log)"
for v in 0 1 0 1; do
  echo "JAT_AGB_STAGED=$v"; JAT_AGB_STAGED=$v timeout 300 python scripts/bench_kernels.py --only bwd --iters 30 2>&1 | grep "adaln_gate_bwd"
done
for v in 0 1 0 1; do
  JAT_AGB_STAGED=$v timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/train_$v.log 2> gpurun_out/train_$v.err || tail -3 gpurun_out/train_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/train_$v.log').read().strip().splitlines()[-1])
t=d.get('train', d)
print('STAGED=$v ms/step', t['ms_per_step'], 'adaln_bwd', t['kernels']['adaln_bwd']['ms_per_step'], 'loss', t['loss'], 'clk', t['clocks']['sm_mhz'])
PY
done
