#!/bin/bash
# Whole GPU suite, smoke, and both benches.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/all_tests.log 2>&1
echo "all_tests exit=$? :: $(tail -n 3 gpurun_out/all_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"; python scripts/show_bench.py gpurun_out/bench.log 2>/dev/null | head -12
timeout 900 python bench.py --mode train --steps 6 --warmup 3 > gpurun_out/bench_train.log 2> gpurun_out/bench_train.err
echo "bench_train exit=$? :: $(head -c 700 gpurun_out/bench_train.log)"
