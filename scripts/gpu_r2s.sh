#!/bin/bash
# N-GPU DDP training step with the final build (default exchange settings)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/r2s_train_n$N.log 2> gpurun_out/r2s_train_n$N.err
echo "bench train N=$N exit=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_train_n$N.log').read().strip().splitlines()[-1])
    print(' ms/step', d['ms_per_step'], 'no_sync', d.get('no_sync_ms_per_step'), 'exposed', d.get('allreduce_exposed_ms'), 'ranks identical', d.get('params_identical_across_ranks'), 'loss', d['loss'], 'clk', d['clocks'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r2s_train_n$N.err').read()[-1500:])
PY
