#!/bin/bash
# round-2 visit A: full parity suite + smoke + bench (all legs) on one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout=1200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -20
tail -3 gpurun_out/smoke.log
tail -5 gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-6000
