#!/bin/bash
# evidence with the tile backward as default: bench line, serialised launch list of the eager step, full ncu capture of
# the tile plan builder + gather
mkdir -p gpurun_out
( time timeout 300 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-200
timeout 300 bash tools/gpu_launches.sh
timeout 200 bash tools/gpu_ncu_roi_bwd_tile.sh 2
