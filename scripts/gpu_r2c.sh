#!/bin/bash
# Round 2: quick kernel iteration -- attention tests + training bench (per-class kernel times).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "${1:-attention}" > gpurun_out/r2c_tests.log 2>&1
echo "tests exit=$? :: $(tail -n 3 gpurun_out/r2c_tests.log)"
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2c_train.log 2> gpurun_out/r2c_train.err
echo "bench train exit=$?"; tail -3 gpurun_out/r2c_train.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_train.log').read().strip().splitlines()[-1])
print(' ms/step', d['ms_per_step'], 'frac', d['step_tensor_frac_sustained'], 'clocks', d['clocks'])
for n,e in sorted(d['kernels'].items(), key=lambda x:-x[1]['ms_per_step'])[:12]: print('   ', n, e['ms_per_step'], 'x', e['launches_per_step'])
PY
