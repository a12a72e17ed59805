#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_lpips_gpu.py tests/test_image_glue_gpu.py tests/test_distillation_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
grep -E "passed|failed|^FAILED|^E  |LPIPS|rel " gpurun_out/pytest_quick.log | cut -c1-250 | tail -30
tail -2 gpurun_out/bench.log | cut -c1-700
