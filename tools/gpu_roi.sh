#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roi_align.py -q -p no:cacheprovider --tb=short -x > gpurun_out/pytest_roi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_roi.log
tail -n 25 gpurun_out/pytest_roi.log
timeout 300 python tools/roi_microbench.py > gpurun_out/roi_micro.log 2>&1
timeout 300 python tools/roi_microbench.py --bwd >> gpurun_out/roi_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/roi_micro.log
cat gpurun_out/roi_micro.log
