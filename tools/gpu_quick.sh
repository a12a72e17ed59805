#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_lpips_gpu.py tests/test_ngp_render_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
grep -E "passed|failed|^FAILED|^E  |LPIPS|C2 view" gpurun_out/pytest_quick.log | cut -c1-250 | tail -20
