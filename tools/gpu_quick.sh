#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-300; grep -o '"cpu_baseline".*' gpurun_out/bench.log | cut -c1-700
