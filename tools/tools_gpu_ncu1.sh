#!/bin/bash
# usage: tools_gpu_ncu1.sh <kernel-regex> <skip> ; full ncu capture of one kernel inside a short bench run
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-3} -c 1 -o gpurun_out/prof_one $BENCH > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
