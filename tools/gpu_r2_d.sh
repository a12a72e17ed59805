#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_minibatch.py > gpurun_out/diag_minibatch.log 2>&1
for m in 0x7fffffff 0x7fffffef 0x7ffffff7; do
  echo "== fusion mask $m" >> gpurun_out/trace_variants.log
  timeout 300 python tools/microbench.py trace fuse=$m 2>&1 | grep -E "trace:|gca|gate_mlp|linear_small|gn_cluster" >> gpurun_out/trace_variants.log
done
tail -12 gpurun_out/diag_minibatch.log
cat gpurun_out/trace_variants.log
