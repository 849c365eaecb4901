#!/bin/bash
# ragged heads-per-CTA attention forward + single-wave colsum: kernel tests, training step, sampling step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x > gpurun_out/k_tests.log 2>&1
echo "kernel_tests exit=$? :: $(tail -n 2 gpurun_out/k_tests.log)"
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/train.log 2> gpurun_out/train.err || tail -3 gpurun_out/train.err
python - <<PY
import json
d=json.loads(open('gpurun_out/train.log').read().strip().splitlines()[-1])
t=d.get('train', d)
print('train ms/step', t['ms_per_step'], 'loss', t['loss'], 'clk', t['clocks']['sm_mhz'])
for k in ('gqa_attention_fwd','colsum_cast','adaln_bwd','attention_bwd'): print('  ',k, t['kernels'][k]['ms_per_step'])
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sub none > gpurun_out/sample.log 2> gpurun_out/sample.err || tail -3 gpurun_out/sample.err
python scripts/show_bench.py gpurun_out/sample.log 2>/dev/null | head -14
