#!/bin/bash
# visit P: large-image (128 x 128 latents) SIMT kernels -- parity then in-graph traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_ops_gpu.py tests/test_unet_gpu.py -x -q -m gpu > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/p_tests.log
tail -5 gpurun_out/p_tests.log
timeout 600 python tools/microbench.py trace hw=128 2>&1 | grep -E "trace:|  " | head -24 > gpurun_out/p_trace_128.log
timeout 600 python tools/microbench.py trace 2>&1 | grep -E "trace:|  " | head -16 >> gpurun_out/p_trace_128.log
timeout 600 python tools/microbench.py unet x3only 2>&1 | grep -E "^unet" >> gpurun_out/p_trace_128.log
cat gpurun_out/p_trace_128.log
