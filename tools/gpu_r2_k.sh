#!/bin/bash
# visit K: new tests (128x128-latent UNet, sliced render, EFT), GroupNorm cluster variants (trace + batch 8/16), C5 leg on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_ngp_render_gpu.py tests/test_unet_ops_gpu.py tests/test_eft_gpu.py tests/test_vae_gpu.py -q --timeout=900 -p no:cacheprovider -rA > gpurun_out/pytest_k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_k.log
rm -f gpurun_out/trace_variants.log
for m in 0x7fffffff 0x7ffffff7 0x7ffffffb; do
  echo "== fusion mask $m (bit2 = cluster GN, bit3 = adaptive cluster width)" >> gpurun_out/trace_variants.log
  timeout 300 python tools/microbench.py trace fuse=$m 2>&1 | grep -E "trace:|gn_" >> gpurun_out/trace_variants.log
  timeout 300 python tools/microbench.py unet x3only nb16 fuse=$m 2>&1 | grep -E "^unet" >> gpurun_out/trace_variants.log
done
for m in 0x7fffffff 0x7fffffef; do
  echo "== render, fusion mask $m (bit4 = tiled field kernels)" >> gpurun_out/trace_variants.log
  timeout 300 python tools/microbench.py render fuse=$m 2>&1 | grep -E "^render" >> gpurun_out/trace_variants.log
done
timeout 1200 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-c4 --no-c2 --no-gpuref --c5 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err
echo "bench c5 exit $?" >> gpurun_out/bench_c5.err
grep -E "passed|failed" gpurun_out/pytest_k.log | tail -3
grep -E "^FAILED|^ERROR|rel vs|128x128|rel " gpurun_out/pytest_k.log | head -20
cat gpurun_out/trace_variants.log
tail -3 gpurun_out/bench_c5.err
tail -1 gpurun_out/bench_c5.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['c5_large_latents'])"
