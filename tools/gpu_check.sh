#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, micro-benchmarks.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -m pytest tests -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
timeout 600 python tools/microbench.py unet render > gpurun_out/microbench.log 2>&1
echo "microbench exit: $?" >> gpurun_out/microbench.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
tail -3 gpurun_out/smoke.log
tail -4 gpurun_out/bench.log | cut -c1-1500
tail -6 gpurun_out/microbench.log
