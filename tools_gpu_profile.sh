#!/bin/bash
# tests + smoke + bench, then the ncu launch list and one full capture of the ROIAlign kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
$BENCH > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd -s 3 -c 1 -o gpurun_out/prof_roi $BENCH > gpurun_out/ncu_full.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log | cut -c1-600; tail -n 3 gpurun_out/ncu_list.log; tail -n 3 gpurun_out/ncu_full.log
