#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_roi_align.py -q -p no:cacheprovider --tb=short > gpurun_out/pytest_roi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_roi.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -n 15 gpurun_out/pytest_roi.log; grep '^{' gpurun_out/bench.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['stage_ms'])"
