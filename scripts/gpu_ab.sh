#!/bin/bash
# A/B of an environment knob on the sampling bench: gpu_ab.sh VAR "v1 v2" [repeats]
mkdir -p gpurun_out
for r in $(seq 1 ${3:-2}); do for v in $2; do
  env $1=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1])
print('$1=$v', 'steps/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'] if d.get('e2e') else None, 'clk', d['clocks']['sm_mhz'])
PY
done; done
