#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_unet_ops_gpu.py -q --timeout=800 -p no:cacheprovider > gpurun_out/pytest_l.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_l.log
timeout 300 python tools/microbench.py trace 2>&1 | grep -E "trace:|  " | head -12 > gpurun_out/trace_l.log
timeout 300 python tools/microbench.py unet x3only nb16 2>&1 | grep -E "^unet" >> gpurun_out/trace_l.log
tail -4 gpurun_out/pytest_l.log; cat gpurun_out/trace_l.log
