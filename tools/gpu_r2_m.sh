#!/bin/bash
# multi-GPU visit: bench.py through torchrun on N GPUs of one box (N = $1), plus the 2-rank NCCL minibatch check on real devices
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multirank_gpu.py tests/test_minibatch_gpu.py -x -q -m gpu > gpurun_out/pytest_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n$N.log; tail -3 gpurun_out/pytest_n$N.log
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/nvidia_smi_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?" >> gpurun_out/bench_n$N.err
tail -4 gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'views_per_step', 'views_per_s', 'optimizer_steps_per_s', 'e2e', 'clocks')})
print('c4', d['c4_fixed_views']); print('c5', d.get('c5_large_latents'))
"
