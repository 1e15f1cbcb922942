#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider --tb=short -x -k "kd_loss or distillation" > gpurun_out/pytest_distill.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_distill.log
tail -n 30 gpurun_out/pytest_distill.log | cut -c1-300
