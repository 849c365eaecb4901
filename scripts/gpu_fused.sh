#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "adaln or gate_bwd or train or dropout or optim" > gpurun_out/fused_tests.log 2>&1
echo "tests exit=$? :: $(tail -n 6 gpurun_out/fused_tests.log)"
timeout 900 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_train.log 2> gpurun_out/bench_train.err
echo "bench_train exit=$?"; tail -3 gpurun_out/bench_train.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_train.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'tf',d['step_tflops_per_gpu'],'frac',d['step_tensor_frac_sustained'],'kernel_ms',d['kernel_ms_per_step'],'loss',d['loss'],d['clocks'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']): print(f"  {k:22s} {v['ms_per_step']:7.3f} ms  x{v['launches_per_step']}")
PY
