#!/bin/bash
# what the driver runs at round end, in its order: gpu tests, smoke, reference arm, own arm (default flags)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.log
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench_ref.log | cut -c1-300; grep real gpurun_out/bench_ref.err; tail -n 2 gpurun_out/bench.log | cut -c1-300; grep real gpurun_out/bench.err
