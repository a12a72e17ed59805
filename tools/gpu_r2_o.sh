#!/bin/bash
# visit O: 256-thread field backward -- parity then timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ngp_render_gpu.py tests/test_distillation_gpu.py -x -q -m gpu > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/o_tests.log
tail -5 gpurun_out/o_tests.log
timeout 300 python tools/microbench.py render > gpurun_out/o_render.log 2>&1; tail -20 gpurun_out/o_render.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-c4 --no-c2 --no-gpuref > gpurun_out/o_bench.log 2>&1; tail -2 gpurun_out/o_bench.log
