#!/bin/bash
# round-2 visit C: batch-row diagnostic, cluster kernels v2 (unrolled memory loops), 128x128-latent UNet tests, in-graph trace
mkdir -p gpurun_out
timeout 600 python tools/diag_batch_rows.py > gpurun_out/diag_batch_rows.log 2>&1
timeout 900 python -m pytest tests/test_unet_ops_gpu.py tests/test_unet_gpu.py -q -x --timeout=800 -p no:cacheprovider -k "gca or groupnorm or 128 or full_unet" > gpurun_out/pytest_c.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_c.log
timeout 600 python tools/microbench.py trace > gpurun_out/microbench_c.log 2>&1
cat gpurun_out/diag_batch_rows.log | tail -40
tail -6 gpurun_out/pytest_c.log
grep -E "^FAILED|^ERROR|Error|rel vs" gpurun_out/pytest_c.log | head
grep -E "trace:|  " gpurun_out/microbench_c.log | head -24
