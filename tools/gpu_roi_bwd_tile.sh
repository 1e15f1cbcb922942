#!/bin/bash
# pixel-tile ROIAlign backward: parity tests, then the backward microbench (all five variants)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_roi_align.py -q -p no:cacheprovider --tb=short -k "bwd" > gpurun_out/pytest_roi_tile.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_roi_tile.log
tail -n 40 gpurun_out/pytest_roi_tile.log
timeout 200 python tools/roi_microbench.py --bwd --iters 10 > gpurun_out/roi_tile_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/roi_tile_micro.log
cat gpurun_out/roi_tile_micro.log
