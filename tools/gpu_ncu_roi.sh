#!/bin/bash
# usage: gpu_ncu_roi.sh <kernel-regex> [impls] [steps] ; full ncu capture of one ROIAlign kernel launch in the microbench
mkdir -p gpurun_out
CMD="python tools/roi_microbench.py --iters 2 --impls ${2:-2} --steps ${3:-1}"
$CMD > gpurun_out/plain_roi.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 1 -f -o gpurun_out/prof_roi $CMD > gpurun_out/ncu_roi.log 2>&1
tail -n 3 gpurun_out/ncu_roi.log | cut -c1-200
