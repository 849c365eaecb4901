#!/bin/bash
# A/B of two builds of the library on the sampling bench and the attention micro-benchmark:
#   gpu_ab_lib.sh <old.so> [repeats]     (new = the in-tree build)
mkdir -p gpurun_out
OLD=$(realpath $1)
for r in $(seq 1 ${2:-2}); do for v in old new; do
  if [ $v = old ]; then export JAT_B200_LIB=$OLD; else unset JAT_B200_LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sub none > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1])
k=d.get('kernels') or {}
print('$v', 'steps/s', d['value'], 'ms', d['ms_per_step'], 'clk', d['clocks']['sm_mhz'], 'attn', (k.get('gqa_attention_fwd') or {}).get('ms_per_step'))
PY
  timeout 300 python scripts/bench_kernels.py --only attn --iters 20 2>&1 | grep -i "attn\|attention" | head -4
done; done
