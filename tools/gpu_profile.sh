#!/bin/bash
# one `ncu --set full` capture of the hot kernels (UNet convs, NGP field kernels).  Reports stay on the box (too big for gpurun_out);
# what comes back: the raw-page CSV of every launch and the gzipped source-page CSV (per-instruction stall samples).
mkdir -p gpurun_out
T=/tmp/prof; mkdir -p $T
timeout 1500 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm -o $T/conv_full -f python tools/profile_targets.py unet 1 > gpurun_out/prof_conv_full.log 2>&1
echo "conv_full exit $?"
ncu -i $T/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2> gpurun_out/conv_full_raw.err
ncu -i $T/conv_full.ncu-rep --page source --csv 2> gpurun_out/conv_full_src.err | gzip -9 > gpurun_out/conv_full_source.csv.gz
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"field_|mlp_wgrad|ray_" -o $T/render_full -f python tools/profile_targets.py render > gpurun_out/prof_render_full.log 2>&1
echo "render_full exit $?"
ncu -i $T/render_full.ncu-rep --page raw --csv > gpurun_out/render_full_raw.csv 2> gpurun_out/render_full_raw.err
ncu -i $T/render_full.ncu-rep --page source --csv 2> gpurun_out/render_full_src.err | gzip -9 > gpurun_out/render_full_source.csv.gz
ls -la $T gpurun_out
S=$(stat -c %s $T/render_full.ncu-rep); if [ "$S" -lt 30000000 ]; then cp $T/render_full.ncu-rep gpurun_out/; fi
du -sh gpurun_out
