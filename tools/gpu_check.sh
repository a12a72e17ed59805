#!/bin/bash
# One GPU-box visit: parity tests, smoke, micro-benchmarks.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -m pytest tests -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python tools/microbench.py unet render > gpurun_out/microbench.log 2>&1
echo "microbench exit: $?" >> gpurun_out/microbench.log
tail -40 gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log
