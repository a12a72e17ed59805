#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_minibatch_rep.log
for i in 1 2 3 4; do timeout 300 python tools/diag_minibatch.py 2>&1 | grep "view " >> gpurun_out/diag_minibatch_rep.log; done
cat gpurun_out/diag_minibatch_rep.log
bash tools/gpu_r2_a.sh
