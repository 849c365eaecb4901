#!/bin/bash
# usage: gpu_quick.sh "<pytest -k expression or empty>"  -> selected tests, then the sampling bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x ${1:+-k "$1"} > gpurun_out/quick_tests.log 2>&1
echo "tests exit=$? :: $(tail -n 4 gpurun_out/quick_tests.log)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"; python scripts/show_bench.py gpurun_out/bench.log 2>/dev/null || tail -c 400 gpurun_out/bench.log
