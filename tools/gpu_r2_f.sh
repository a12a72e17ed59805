#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eft_gpu.py tests/test_unet_ops_gpu.py -q --timeout=500 -p no:cacheprovider > gpurun_out/pytest_f.log 2>&1
for i in 1 2 3; do timeout 300 python tools/diag_minibatch.py 2>&1 | grep "view " >> gpurun_out/diag_minibatch_rep.log; done
timeout 900 python -m pytest tests/test_minibatch_gpu.py tests/test_multirank_gpu.py tests/test_distillation_gpu.py -q --timeout=800 -p no:cacheprovider > gpurun_out/pytest_f2.log 2>&1
tail -30 gpurun_out/pytest_f.log | grep -v "^$" | tail -25
cat gpurun_out/diag_minibatch_rep.log
tail -8 gpurun_out/pytest_f2.log
