#!/bin/bash
# round-2 ncu captures (one GPU, --clock-control none; reports stay on the box, raw-page CSVs come back):
#   conv_b1   : the conv / linear launches of one batch-1 UNet evaluation            (weight-streaming regime: HBM %)
#   conv_b16  : the same launches at batch 16 (the view-batched step's UNet batch)   (tensor-bound regime: tensor-pipe %)
#   simt_b1   : GroupNorm (cluster), GlobalContext, LayerNorm, attention, ... at batch 1
#   render    : NGP field / per-ray kernels of one 128x128 render forward + backward
#   tensor metrics: an explicit --metrics pass listing every tensor-pipe counter this ncu knows for the device
mkdir -p gpurun_out
T=/tmp/prof; mkdir -p $T
ncu --query-metrics 2>/dev/null | grep -i -E "pipe_tensor|tcgen|utc" | awk '{print $1}' | sort -u > gpurun_out/tensor_metric_names.txt
MET=$(grep -E "^sm__pipe_tensor.*cycles_active$|^sm__inst_executed_pipe_tensor" gpurun_out/tensor_metric_names.txt | head -12 | sed 's/$/.avg.pct_of_peak_sustained_active/' | grep cycles_active | paste -sd, -)
echo "tensor metrics: $MET" > gpurun_out/prof_r2.log
for what in "conv_b1 1 conv_gemm" "conv_b16 16 conv_gemm"; do
  set -- $what
  timeout 1500 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$3 -o $T/$1 -f python tools/profile_targets.py unet $2 >> gpurun_out/prof_r2.log 2>&1
  echo "$1 exit $?" >> gpurun_out/prof_r2.log
  ncu -i $T/$1.ncu-rep --page raw --csv > gpurun_out/${1}_raw.csv 2> gpurun_out/${1}_raw.err
  if [ -n "$MET" ]; then
    timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,$MET --clock-control none -k regex:$3 --csv --log-file gpurun_out/${1}_tensor.csv python tools/profile_targets.py unet $2 >> gpurun_out/prof_r2.log 2>&1
  fi
done
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"gn_cluster|gn_apply|gn_stats|gca_|gate_mlp|linear_small|layernorm_rows|mq_attention|cross_attention|concat2|pixel_shuffle" -o $T/simt_b1 -f python tools/profile_targets.py unet 1 >> gpurun_out/prof_r2.log 2>&1
echo "simt_b1 exit $?" >> gpurun_out/prof_r2.log
ncu -i $T/simt_b1.ncu-rep --page raw --csv > gpurun_out/simt_b1_raw.csv 2> gpurun_out/simt_b1_raw.err
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"field_|mlp_wgrad|ray_|tape_rowsum" -o $T/render -f python tools/profile_targets.py render >> gpurun_out/prof_r2.log 2>&1
echo "render exit $?" >> gpurun_out/prof_r2.log
ncu -i $T/render.ncu-rep --page raw --csv > gpurun_out/render_raw.csv 2> gpurun_out/render_raw.err
# launch list of the bench command itself (cold-cache, serialised: shares only)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-c4 --no-c2 --no-gpuref > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?" >> gpurun_out/prof_r2.log
gzip -9 -f gpurun_out/launches_bench.csv
tail -12 gpurun_out/prof_r2.log
wc -l gpurun_out/*_raw.csv gpurun_out/*_tensor.csv 2>/dev/null
head -c 600 gpurun_out/tensor_metric_names.txt
du -sh gpurun_out
