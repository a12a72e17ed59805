#!/bin/bash
# visit R: ncu --set full of the render kernels (64-point-tile backward), raw + source pages for the backward kernel
mkdir -p gpurun_out
T=/tmp/prof; mkdir -p $T
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"field_|ray_" -o $T/render -f python tools/profile_targets.py render > gpurun_out/r_prof.log 2>&1
echo "render exit $?" >> gpurun_out/r_prof.log
ncu -i $T/render.ncu-rep --page raw --csv > gpurun_out/render_raw.csv 2> gpurun_out/render_raw.err
ncu -i $T/render.ncu-rep --page details --csv -k regex:field_backward > gpurun_out/render_bwd_details.csv 2>/dev/null
ncu -i $T/render.ncu-rep --page source --csv -k regex:field_backward > gpurun_out/render_bwd_source.csv 2>/dev/null
tail -3 gpurun_out/r_prof.log; wc -l gpurun_out/render_*.csv
