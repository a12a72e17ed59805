#!/bin/bash
mkdir -p gpurun_out
for v in "" "nocarve" "nopdl"; do
  timeout 300 python tools/microbench.py trace $v > gpurun_out/trace_${v:-default}.log 2>&1
  echo "== trace $v"; head -12 gpurun_out/trace_${v:-default}.log
done
timeout 300 python tools/microbench.py phases > gpurun_out/conv_phases.log 2>&1; cat gpurun_out/conv_phases.log | tail -20
