#!/bin/bash
# final visit: full parity suite, smoke, render capture of the tiled kernels, both bench arms
mkdir -p gpurun_out
T=/tmp/prof; mkdir -p $T
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=1200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"field_|ray_" -o $T/render_v2 -f python tools/profile_targets.py render > gpurun_out/prof_render_v2.log 2>&1
ncu -i $T/render_v2.ncu-rep --page raw --csv > gpurun_out/render_v2_raw.csv 2> gpurun_out/render_v2_raw.err
timeout 700 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --clock-control none -k regex:"gn_cluster|gn_apply|gn_stats|gca_|gate_mlp|linear_small|layernorm_rows|mq_attention|cross_attention|concat2|pixel_shuffle|im2col|nchw|nhwc" -o $T/simt_b1 -f python tools/profile_targets.py unet 1 > gpurun_out/prof_simt_b1.log 2>&1
ncu -i $T/simt_b1.ncu-rep --page raw --csv > gpurun_out/simt_b1_raw.csv 2> gpurun_out/simt_b1_raw.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
echo "bench reference exit: $?" >> gpurun_out/bench_reference.err
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
tail -2 gpurun_out/smoke.log
tail -2 gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-7000
tail -1 gpurun_out/bench_reference.log | cut -c1-1800
