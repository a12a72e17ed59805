#!/bin/bash
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 --launch-timeout 0 python tools/diag_min_once.py 2 nograph > gpurun_out/sanitizer_memcheck_nograph.log 2>&1
echo "memcheck nograph exit $?"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 python tools/diag_min_once.py 2 > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"
timeout 1200 compute-sanitizer --tool initcheck --print-limit 30 python tools/diag_min_once.py 2 nograph > gpurun_out/sanitizer_initcheck.log 2>&1
echo "initcheck exit $?"
for f in gpurun_out/sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|Invalid|Uninitialized|at .* in |finite per view|by thread|Address" $f | head -40; done
