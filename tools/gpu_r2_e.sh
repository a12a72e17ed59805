#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/diag_uninit.py > gpurun_out/diag_uninit.log 2>&1
timeout 600 python -m pytest tests/test_eft_gpu.py -q -x --timeout=500 -p no:cacheprovider > gpurun_out/pytest_eft.log 2>&1
cat gpurun_out/diag_uninit.log | tail -40
tail -25 gpurun_out/pytest_eft.log
