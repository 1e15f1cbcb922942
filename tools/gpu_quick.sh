#!/bin/bash
# all gpu tests + the default bench (no CPU baseline): quick check after a kernel change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
echo "bench exit $?" >> gpurun_out/bench_quick.log
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300; tail -n 2 gpurun_out/bench_quick.log | cut -c1-200; tail -n 3 gpurun_out/bench_quick.err
