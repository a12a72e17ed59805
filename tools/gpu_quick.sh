#!/bin/bash
# quick visit: selected tests + bench + unet microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_v2_gpu.py tests/test_conv_gpu.py tests/test_unet_gpu.py tests/test_distillation_gpu.py tests/test_ngp_render_gpu.py -m gpu -q -rA --timeout=600 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
timeout 600 python tools/microbench.py unet > gpurun_out/microbench.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_quick.log | cut -c1-200 | tail -25
tail -4 gpurun_out/bench.log | cut -c1-3000
tail -5 gpurun_out/microbench.log
