#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_vae_gpu.py tests/test_unet_ops_gpu.py tests/test_distillation_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 600 python tools/microbench.py vaetrace > gpurun_out/vaetrace.log 2>&1
grep -E "passed|failed|^FAILED|^E  |rel " gpurun_out/pytest_quick.log | cut -c1-200 | tail -12
cat gpurun_out/vaetrace.log | cut -c1-600
