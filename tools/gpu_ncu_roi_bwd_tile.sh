#!/bin/bash
# full ncu capture of the pixel-tile backward's plan builder and gather (bin_step ${1:-2}) in the microbench
mkdir -p gpurun_out
CMD="python tools/roi_microbench.py --bwd --iters 1 --steps ${1:-2} --variants 3"
ncu --set full --clock-control none --import-source on -k regex:roi_bwd_tile -c 2 -f -o gpurun_out/prof_roi_bwd_tile $CMD > gpurun_out/ncu_roi_bwd_tile.log 2>&1
tail -n 3 gpurun_out/ncu_roi_bwd_tile.log | cut -c1-200
