#!/bin/bash
# visit T: per-kernel trace of one VAE decode / encode (256 x 256)
mkdir -p gpurun_out
timeout 600 python tools/microbench.py vaetrace vae > gpurun_out/t_vae.log 2>&1; grep -v "^$" gpurun_out/t_vae.log | cut -c1-400 | tail -60
