#!/bin/bash
# ncu source-level captures (<= 3 reports per call: gpurun_out/ must stay under 64 MiB)
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'.'); from jat_b200 import _lib; _lib.load()" || exit 1
timeout 300 python scripts/prof_train_gemms.py > gpurun_out/r2e_gemms_plain.log 2>&1 &&
timeout 300 python scripts/attb_trace.py > gpurun_out/r2e_attb_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gqa_attention_bwd_kernel -s 3 -c 1 -o gpurun_out/r2e_attb python scripts/attb_trace.py > gpurun_out/r2e_ncu_attb.log 2>&1
echo "ncu attb exit=$?"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r2e_fc1_train python scripts/prof_train_gemms.py > gpurun_out/r2e_ncu_fc1.log 2>&1
echo "ncu fc1 exit=$?"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 20 -c 1 -o gpurun_out/r2e_dact python scripts/prof_train_gemms.py > gpurun_out/r2e_ncu_dact.log 2>&1
echo "ncu dact exit=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
