#!/bin/bash
# Round 2, first GPU pass: the new parity / protocol / DDP tests, the whole GPU suite, smoke, and the default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests/test_protocol_gpu.py tests/test_ddp_gpu.py tests/test_fullsize_train_parity_gpu.py \
    "tests/test_fullsize_parity_gpu.py::test_c2_v2_full_depth_b1_25_step_sampler" -m gpu -q -x -s > gpurun_out/r2_new_tests.log 2>&1
echo "new_tests exit=$? :: $(tail -n 4 gpurun_out/r2_new_tests.log)"
grep -h "FULLSIZE_TRAIN_PARITY\|C2_PARITY\|DDP_WORKER_OK" gpurun_out/r2_new_tests.log | cut -c1-600
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1
echo "all_tests exit=$? :: $(tail -n 3 gpurun_out/all_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"; python scripts/show_bench.py gpurun_out/bench.log 2>/dev/null | head -40
tail -5 gpurun_out/bench.err
