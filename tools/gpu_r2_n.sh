#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/microbench.py trace hw=128 2>&1 | grep -E "trace:|  " | head -24 > gpurun_out/trace_128.log
timeout 600 python tools/microbench.py unet x3only nb32 2>&1 | grep -E "^unet" >> gpurun_out/trace_128.log
cat gpurun_out/trace_128.log
