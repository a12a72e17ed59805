#!/bin/bash
# final visit of the round: full parity suite, smoke, both bench arms (C5 leg included), launch list of the bench command under ncu
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout=1200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --c5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
echo "bench reference exit: $?" >> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-c4 --no-c2 --no-gpuref > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?" >> gpurun_out/bench.err
gzip -9 -f gpurun_out/launches_bench.csv
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
tail -2 gpurun_out/smoke.log
tail -3 gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-8000
tail -1 gpurun_out/bench_reference.log | cut -c1-1800
