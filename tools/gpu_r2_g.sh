#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_minibatch2.py > gpurun_out/diag_minibatch2.log 2>&1
tail -20 gpurun_out/diag_minibatch2.log
