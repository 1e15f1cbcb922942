#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_head.py -q -p no:cacheprovider --tb=short -x > gpurun_out/pytest_train.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -n 40 gpurun_out/pytest_train.log
