#!/bin/bash
# last check of the round: ROIAlign backward tests, smoke(), default bench line
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_roi_align.py -q -p no:cacheprovider --tb=short -k "bwd" 2>&1 | tail -n 2
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
( time timeout 200 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-160; grep real gpurun_out/bench.err
