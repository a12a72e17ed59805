#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_v2_gpu.py tests/test_unet_gpu.py tests/test_distillation_gpu.py -m gpu -q -rA --timeout=600 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 600 python tools/microbench.py unet vae > gpurun_out/microbench.log 2>&1
timeout 600 python tools/microbench.py unet nocarve > gpurun_out/microbench_nocarve.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
grep -E "passed|failed|^FAILED|^E  |loss A|rel diff|PLMS" gpurun_out/pytest_quick.log | cut -c1-200 | tail -25
tail -6 gpurun_out/microbench.log; echo nocarve; tail -4 gpurun_out/microbench_nocarve.log
tail -2 gpurun_out/bench.log | cut -c1-2500
