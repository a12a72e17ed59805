#!/bin/bash
mkdir -p gpurun_out
T=/tmp/prof; mkdir -p $T
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"gn_grid|gn_apply|gn_stats|gca_|gate_mlp|linear_small|layernorm_rows|mq_attention|cross_attention|concat2|pixel_shuffle" -o $T/simt_full -f python tools/profile_targets.py unet 1 > gpurun_out/prof_simt_full.log 2>&1
echo "simt_full exit $?"
ncu -i $T/simt_full.ncu-rep --page raw --csv > gpurun_out/simt_full_raw.csv 2> gpurun_out/simt_full_raw.err
ls -la $T; wc -l gpurun_out/simt_full_raw.csv
