#!/bin/bash
# visit Q: 64-point-tile field backward (2 CTAs/SM) + batch-32 view chunks -- parity, then timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ngp_render_gpu.py tests/test_distillation_gpu.py tests/test_minibatch_gpu.py -x -q -m gpu > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/q_tests.log
tail -5 gpurun_out/q_tests.log
timeout 300 python tools/microbench.py render > gpurun_out/q_render.log 2>&1; tail -3 gpurun_out/q_render.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c2 --no-gpuref > gpurun_out/q_bench.log 2>&1; tail -1 gpurun_out/q_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','c4_fixed_views')})"
