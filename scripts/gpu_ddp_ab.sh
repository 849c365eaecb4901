#!/bin/bash
# usage: gpu_ddp_ab.sh N  -> DDP training bench on N GPUs under a few overlap settings (bucket size / bucket views / SM reserve)
N=$1
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 6 --warmup 3 > gpurun_out/ddp_$name.log 2> gpurun_out/ddp_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ddp_$name.log').read().strip().splitlines()[-1])
    print('$name', 'rc=$rc', 'steps/s', d['value'], 'ms', d['ms_per_step'], 'kernel_ms', d['kernel_ms_per_step'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$name', 'rc=$rc', 'FAILED', e)
PY
}
run base JAT_X=0
run view JAT_DDP_BUCKET_VIEW=1
run b200 JAT_DDP_BUCKET_VIEW=1 JAT_DDP_BUCKET_MB=200
run r8 JAT_DDP_BUCKET_VIEW=1 JAT_SM_RESERVE=8
run r16 JAT_DDP_BUCKET_VIEW=1 JAT_SM_RESERVE=16
run r16b200 JAT_DDP_BUCKET_VIEW=1 JAT_DDP_BUCKET_MB=200 JAT_SM_RESERVE=16
run r16ch8 JAT_DDP_BUCKET_VIEW=1 JAT_SM_RESERVE=16 NCCL_MAX_NCHANNELS=8
tail -n 2 gpurun_out/ddp_base.err
