#!/bin/bash
# usage: gpu_ddp8.sh N -> DDP training bench on N GPUs: default (f32 all-reduce) and with the bf16 compression hook
N=$1
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 6 --warmup 3 > gpurun_out/bench_train_n${N}_$name.log 2> gpurun_out/bench_train_n${N}_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_train_n${N}_$name.log').read().strip().splitlines()[-1])
    print('N=$N $name', 'rc=$rc', 'steps/s', d['value'], 'ms', d['ms_per_step'], 'kernel_ms', d['kernel_ms_per_step'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('N=$N $name', 'rc=$rc', 'FAILED', e)
PY
}
run ${2:-f32} ${3:-JAT_X=0}
# (a bf16_compress_hook variant was measured once: 74.0 ms / step at 8 GPUs, slower than the f32 all-reduce)
