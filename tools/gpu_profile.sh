#!/bin/bash
# launch lists (per-kernel device time) for one UNet forward (B=1) and one render fwd+bwd
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet_b1.csv python tools/profile_targets.py unet 1 > gpurun_out/prof_unet.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_render.csv python tools/profile_targets.py render > gpurun_out/prof_render.log 2>&1
python -m pytest tests/test_ngp_render_gpu.py -m gpu -q -rA --timeout=600 -p no:cacheprovider > gpurun_out/pytest_ngp.log 2>&1
tail -30 gpurun_out/pytest_ngp.log
