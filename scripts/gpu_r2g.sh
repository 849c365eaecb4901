#!/bin/bash
# N-GPU pass: DDP correctness over NCCL + the DDP training step with the f32 and the bf16 gradient exchange
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x -s > gpurun_out/r2g_ddp_tests.log 2>&1
echo "ddp tests exit=$? :: $(tail -n 2 gpurun_out/r2g_ddp_tests.log)"; grep -h "DDP_WORKER_OK" gpurun_out/r2g_ddp_tests.log | cut -c1-330
for wire in f32 bf16; do
  JAT_DDP_GRAD_DTYPE=$wire timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/r2g_train_n${N}_$wire.log 2> gpurun_out/r2g_train_n${N}_$wire.err
  echo "bench train N=$N wire=$wire exit=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2g_train_n${N}_$wire.log').read().strip().splitlines()[-1])
    print(' ms/step', d['ms_per_step'], 'no_sync', d.get('no_sync_ms_per_step'), 'exposed', d.get('allreduce_exposed_ms'), 'ranks identical', d.get('params_identical_across_ranks'), 'loss', d['loss'])
    for n,e in sorted(d['kernels'].items(), key=lambda x:-x[1]['ms_per_step'])[:6]: print('   ', n, e['ms_per_step'], 'x', e['launches_per_step'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r2g_train_n${N}_$wire.err').read()[-1500:])
PY
done
