#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_vae_gpu.py tests/test_unet_gpu.py tests/test_distillation_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 600 python tools/microbench.py vae > gpurun_out/microbench.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
grep -E "passed|failed|^FAILED|^E  |rel " gpurun_out/pytest_quick.log | cut -c1-200 | tail -25
tail -6 gpurun_out/microbench.log
tail -2 gpurun_out/bench.log | cut -c1-3500
