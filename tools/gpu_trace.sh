#!/bin/bash
mkdir -p gpurun_out
for v in "" "nopdl" "nofuse"; do
  timeout 300 python tools/microbench.py trace $v > gpurun_out/trace_${v:-default}.log 2>&1
  echo "== trace $v"; head -14 gpurun_out/trace_${v:-default}.log
done
bash tools/gpu_profile.sh
