#!/bin/bash
# N-GPU sanity of the driver's launch line only (no reference arm, no N=1 rerun)
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_n$N.log 2>&1
echo "bench n=$N exit: $?" >> gpurun_out/bench_n$N.log
grep -E '^\{|exit' gpurun_out/bench_n$N.log | cut -c1-400; tail -3 gpurun_out/bench_n$N.log | cut -c1-200
