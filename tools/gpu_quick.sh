#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multirank_gpu.py tests/test_conv_v2_gpu.py tests/test_unet_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 300 python tools/microbench.py trace > gpurun_out/trace_default.log 2>&1
grep -E "passed|failed|^FAILED|^E  |rel |skip" gpurun_out/pytest_quick.log | cut -c1-250 | tail -20
head -24 gpurun_out/trace_default.log
