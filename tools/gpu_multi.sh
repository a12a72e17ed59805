#!/bin/bash
# N-GPU validation of the driver's launch line (one rank per GPU over NCCL) + the reference arm's rank handling
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/nvidia_smi_multi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_n$N.log 2>&1
echo "bench n=$N exit: $?" >> gpurun_out/bench_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "bench ref n=$N exit: $?" >> gpurun_out/bench_ref_n$N.log
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_n1.log 2>&1
grep -E '^\{|exit' gpurun_out/bench_n$N.log | cut -c1-900
grep -E '^\{|exit' gpurun_out/bench_ref_n$N.log | cut -c1-300
grep -E '^\{|exit' gpurun_out/bench_n1.log | cut -c1-300
tail -5 gpurun_out/bench_n$N.log | cut -c1-300
