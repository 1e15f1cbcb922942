#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_roi_align.py -q -p no:cacheprovider --tb=short -k "tile_gather" 2>&1 | tail -n 3
timeout 40 python tools/roi_microbench.py --bwd --iters 20 --variants 4 --tile-variants 0,3 2>&1 | tee gpurun_out/roi_tile_order.log
