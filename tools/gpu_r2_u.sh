#!/bin/bash
# visit U: column-fixed GroupNorm apply pass, cluster GroupNorm only for <= 1 Mi elements -- parity, then VAE / batch-16 / 128x128 traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_ops_gpu.py tests/test_vae_gpu.py tests/test_unet_gpu.py -x -q -m gpu > gpurun_out/u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/u_tests.log
tail -4 gpurun_out/u_tests.log
timeout 600 python tools/microbench.py vaetrace vae 2>&1 | grep -v "longest" | cut -c1-200 > gpurun_out/u_vae.log; cat gpurun_out/u_vae.log
timeout 600 python tools/microbench.py trace hw=128 2>&1 | grep -E "trace:|  " | head -12 > gpurun_out/u_trace_128.log
timeout 600 python tools/microbench.py unet x3only nb16 2>&1 | grep -E "^unet" >> gpurun_out/u_trace_128.log
cat gpurun_out/u_trace_128.log
