#!/bin/bash
# optimizer-step check: FusedAdamW parity tests, then the training bench with the fused and the torch parameter update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_train_gpu.py -m gpu -q -x > gpurun_out/optim_tests.log 2>&1
echo "tests exit=$? :: $(tail -n 15 gpurun_out/optim_tests.log)"
for v in 0 1; do
JAT_BENCH_TORCH_OPT=$v timeout 900 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_train_opt$v.log 2> gpurun_out/bench_train_opt$v.err
echo "bench_train torch_opt=$v exit=$?"; tail -3 gpurun_out/bench_train_opt$v.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_train_opt$v.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'tf',d['step_tflops_per_gpu'],'frac',d['step_tensor_frac_sustained'],'kernel_ms',d['kernel_ms_per_step'],'loss',d['loss'],d['clocks'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:6]: print(f"  {k:22s} {v['ms_per_step']:7.3f} ms  x{v['launches_per_step']}")
print('  optimizer', d['kernels'].get('optimizer'))
PY
done
