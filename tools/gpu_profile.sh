#!/bin/bash
# one `ncu --set full` capture of the hot kernels (UNet convs, NGP field kernels); read back with `ncu -i ... --page raw --csv`
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:conv_gemm -o gpurun_out/conv_full -f python tools/profile_targets.py unet 1 > gpurun_out/prof_conv_full.log 2>&1
echo "conv_full exit $?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"field_|mlp_wgrad|ray_" -o gpurun_out/render_full -f python tools/profile_targets.py render > gpurun_out/prof_render_full.log 2>&1
echo "render_full exit $?"
ls -la gpurun_out/*.ncu-rep
