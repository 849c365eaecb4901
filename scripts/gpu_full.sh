#!/bin/bash
# Full GPU check: model/sampler parity tests, smoke, short bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x > gpurun_out/model_tests.log 2>&1
echo "model_tests exit=$? :: $(tail -n 1 gpurun_out/model_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$? :: $(tail -c 600 gpurun_out/bench.log)"
tail -n 5 gpurun_out/bench.err
