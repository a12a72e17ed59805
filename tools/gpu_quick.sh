#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_conv_v2_gpu.py tests/test_unet_ops_gpu.py tests/test_vae_gpu.py tests/test_unet_gpu.py tests/test_ngp_render_gpu.py tests/test_distillation_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 600 python tools/microbench.py unet vae > gpurun_out/microbench.log 2>&1
timeout 300 python tools/microbench.py trace > gpurun_out/trace_default.log 2>&1
timeout 300 python tools/microbench.py vaetrace > gpurun_out/vaetrace.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_quick.log | cut -c1-200 | tail -25
tail -7 gpurun_out/microbench.log; head -22 gpurun_out/trace_default.log; grep -E "vae |us " gpurun_out/vaetrace.log | head -14
tail -2 gpurun_out/bench.log | cut -c1-600
