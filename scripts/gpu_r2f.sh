#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/all_tests.log 2>&1
echo "all_tests exit=$? :: $(tail -n 3 gpurun_out/all_tests.log)"
timeout 300 python scripts/attb_trace.py 2>&1 | head -4
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"; python scripts/show_bench.py gpurun_out/bench.log 2>/dev/null | head -70
tail -5 gpurun_out/bench.err
