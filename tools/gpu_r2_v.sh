#!/bin/bash
# visit V: GroupNorm apply pass with rows requested ahead of the statistics fold -- parity, then VAE trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_ops_gpu.py tests/test_vae_gpu.py -x -q -m gpu > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_tests.log
tail -4 gpurun_out/v_tests.log
timeout 600 python tools/microbench.py vaetrace vae 2>&1 | grep -v "longest" | cut -c1-200 > gpurun_out/v_vae.log; cat gpurun_out/v_vae.log
timeout 600 python tools/microbench.py unet x3only nb16 2>&1 | grep -E "^unet" > gpurun_out/v_unet.log; cat gpurun_out/v_unet.log
