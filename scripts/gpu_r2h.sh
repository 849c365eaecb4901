#!/bin/bash
# DDP contention experiments: cap NCCL's CTA count and keep as many SMs out of the persistent GEMM grids
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/r2h_n${N}_$tag.log 2> gpurun_out/r2h_n${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_n${N}_$tag.log').read().strip().splitlines()[-1])
    k=d['kernels']
    print('N=$N $tag: ms/step', d['ms_per_step'], 'host', d.get('host_issue_ms_per_step'), 'no_sync', d.get('no_sync_ms_per_step'), 'exposed', d.get('allreduce_exposed_ms'), 'accum', k['gemm_accum']['ms_per_step'], 'dact', k['gemm_dact']['ms_per_step'], 'bias_act', k['gemm_bias_act']['ms_per_step'], 'kernel_sum', d['kernel_ms_per_step'])
except Exception as e:
    print('$tag parse failed', e); print(open('gpurun_out/r2h_n${N}_$tag.err').read()[-800:])
PY
}
shift
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  run "$tag" $cfg
done
