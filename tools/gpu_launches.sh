#!/bin/bash
# launch list (device time per kernel, serialised) of eager fine-tune steps through forward()
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1500 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
tail -n 2 gpurun_out/ncu_list.log | cut -c1-200
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.md 2> gpurun_out/launches.err; tail -n 3 gpurun_out/launches.md; tail -n 3 gpurun_out/launches.err
