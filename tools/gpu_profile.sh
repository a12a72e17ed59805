#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet_b1.csv python tools/profile_targets.py unet 1 > gpurun_out/prof_unet.log 2>&1
python -m pytest tests/test_conv_v2_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -5
