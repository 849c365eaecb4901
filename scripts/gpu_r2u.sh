#!/bin/bash
# default bench line of the final build (for profiles/)
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
echo "bench exit=$?"; python scripts/show_bench.py gpurun_out/bench_final.log 2>/dev/null | grep -v "^GPU_BASELINE\|^cpu\|^roofline" | head -40 | cut -c1-400
