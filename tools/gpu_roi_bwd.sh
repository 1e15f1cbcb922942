#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi_align.py -q -p no:cacheprovider --tb=short -x -k bwd > gpurun_out/pytest_roi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_roi.log
tail -n 25 gpurun_out/pytest_roi.log
timeout 300 python tools/roi_microbench.py --bwd > gpurun_out/roi_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/roi_micro.log
cat gpurun_out/roi_micro.log
ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:roi_ --csv --log-file gpurun_out/roi_bwd_launches.csv python tools/roi_microbench.py --bwd --iters 1 --steps 2 > gpurun_out/ncu_roi_bwd.log 2>&1
