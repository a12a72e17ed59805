#!/bin/bash
# One GPU-box visit: full parity suite, smoke, both bench arms, micro-benchmarks, ncu launch list of the bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -m pytest tests -m gpu -q -rA --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
echo "bench reference exit: $?" >> gpurun_out/bench_reference.log
timeout 600 python tools/microbench.py unet render vae > gpurun_out/microbench.log 2>&1
timeout 300 python tools/microbench.py trace > gpurun_out/trace_default.log 2>&1
# launch list of the bench command itself (cold-cache, serialised: compare shares, not absolutes; never a bench value)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit: $?" >> gpurun_out/bench_under_ncu.log
gzip -9 -f gpurun_out/launches_bench.csv
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
tail -3 gpurun_out/smoke.log
tail -2 gpurun_out/bench.log | cut -c1-3800
tail -2 gpurun_out/bench_reference.log | cut -c1-1500
tail -8 gpurun_out/microbench.log
ls -la gpurun_out | tail -30
