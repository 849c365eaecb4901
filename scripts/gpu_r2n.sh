#!/bin/bash
# ncu source-level capture of the staged adaln_gate_bwd at the training shape
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adaln_gate_bwd_staged -c 4 -f -o gpurun_out/agb_staged \
  python scripts/bench_kernels.py --only bwd --iters 1 > gpurun_out/agb_ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/agb_ncu.log; ls -la gpurun_out/*.ncu-rep
