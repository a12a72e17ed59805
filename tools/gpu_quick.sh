#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ngp_render_gpu.py tests/test_distillation_gpu.py tests/test_ref_cuda_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 300 python tools/microbench.py render > gpurun_out/microbench.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_quick.log | cut -c1-250 | tail -12
tail -3 gpurun_out/microbench.log
tail -2 gpurun_out/bench.log | cut -c1-700
