#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py render > gpurun_out/microbench.log 2>&1; tail -2 gpurun_out/microbench.log
timeout 600 python -m pytest tests/test_ngp_render_gpu.py tests/test_grid_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -2
