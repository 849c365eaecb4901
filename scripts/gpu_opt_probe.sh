#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/opt_probe.py > gpurun_out/opt_probe.log 2>&1; echo "exit=$?"; cat gpurun_out/opt_probe.log | tail -20
JAT_ADAMW_F32=1 timeout 600 python scripts/opt_probe.py > gpurun_out/opt_probe_f32.log 2>&1; echo "f32 exit=$?"; grep -i "FusedAdamW\|per class" gpurun_out/opt_probe_f32.log
