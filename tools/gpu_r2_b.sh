#!/bin/bash
# round-2 visit B: cluster GroupNorm / GlobalContext kernels -- op tests, UNet parity, in-graph trace, minibatch diagnostic, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_ops_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -q -x --timeout=600 -p no:cacheprovider > gpurun_out/pytest_b.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_b.log
timeout 300 python tools/diag_minibatch.py > gpurun_out/diag_minibatch.log 2>&1
timeout 600 python tools/microbench.py unet trace > gpurun_out/microbench_b.log 2>&1
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err
echo "bench exit: $?" >> gpurun_out/bench_b.err
tail -5 gpurun_out/pytest_b.log
grep -E "^FAILED|^ERROR|Error" gpurun_out/pytest_b.log | head
cat gpurun_out/diag_minibatch.log | tail -30
grep -E "^unet|trace:|  " gpurun_out/microbench_b.log | head -40
tail -3 gpurun_out/bench_b.err
tail -1 gpurun_out/bench_b.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e')})
print('unet_eval_ms', d['roofline']['unet_eval_ms_in_timed_region'], 'frac', d['roofline']['frac'], 'in_graph', d['roofline']['in_graph'])
print('c4', d['c4_fixed_views']); print('gpuref', d['gpu_reference'])
"
