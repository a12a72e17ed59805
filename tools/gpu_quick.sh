#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_image_glue_gpu.py tests/test_multirank_gpu.py tests/test_distillation_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log


grep -E "passed|failed|^FAILED|^E  |rel |skip" gpurun_out/pytest_quick.log | cut -c1-250 | tail -30

