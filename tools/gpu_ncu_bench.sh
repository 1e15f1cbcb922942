#!/bin/bash
# usage: gpu_ncu_bench.sh <kernel-regex> [skip] [extra bench args] ; full ncu capture of one kernel launch inside a short bench run
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${3:-}"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-3} -c 1 -f -o gpurun_out/prof_one $BENCH > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
