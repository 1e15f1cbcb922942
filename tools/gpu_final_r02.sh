#!/bin/bash
# round-end flow with the tile backward as default: gpu tests, smoke, default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 400 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log | cut -c1-600; grep real gpurun_out/bench.err
