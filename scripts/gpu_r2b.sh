#!/bin/bash
# Round 2: attention backward v2 -- parity first, then the training step A/B against v1 (JAT_ATTN_BWD=1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r2b_attn_tests.log 2>&1
echo "attn tests exit=$? :: $(tail -n 3 gpurun_out/r2b_attn_tests.log)"
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_dropout_gpu.py tests/test_fullsize_train_parity_gpu.py -m gpu -q -x -s > gpurun_out/r2b_train_tests.log 2>&1
echo "train tests exit=$? :: $(tail -n 3 gpurun_out/r2b_train_tests.log)"
grep -h "FULLSIZE_TRAIN_PARITY" gpurun_out/r2b_train_tests.log | cut -c1-400
for v in 2 1; do
  JAT_ATTN_BWD=$v timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2b_train_v$v.log 2> gpurun_out/r2b_train_v$v.err
  echo "bench train attn_bwd v$v exit=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2b_train_v$v.log').read().strip().splitlines()[-1])
print(' ms/step', d['ms_per_step'], 'frac', d['step_tensor_frac_sustained'], 'clocks', d['clocks'])
for n,e in sorted(d['kernels'].items(), key=lambda x:-x[1]['ms_per_step'])[:8]: print('   ', n, e['ms_per_step'], 'x', e['launches_per_step'])
PY
done
