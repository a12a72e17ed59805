#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_v2_gpu.py tests/test_unet_gpu.py -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_quick.log
for v in "" fuseall fusenone; do
timeout 300 python tools/microbench.py trace $v > gpurun_out/trace_${v:-default}.log 2>&1
echo "== $v"; head -9 gpurun_out/trace_${v:-default}.log
done
timeout 900 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 --no-cpu > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_quick.log | cut -c1-200 | tail -10
tail -2 gpurun_out/bench.log | cut -c1-600
