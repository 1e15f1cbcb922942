#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 220 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
$BENCH > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"roi_align_fwd|nms_class|text_attention" -s 9 -c 3 -o gpurun_out/prof_top3 $BENCH > gpurun_out/ncu_full.log 2>&1
tail -n 12 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log | cut -c1-3000; tail -n 2 gpurun_out/ncu_list.log | cut -c1-200; tail -n 3 gpurun_out/ncu_full.log | cut -c1-200
