#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fusion.py -q -p no:cacheprovider --tb=short -x -k "teacher or class_mean" > gpurun_out/pytest_teacher.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_teacher.log
timeout 300 python tools/teacher_microbench.py > gpurun_out/teacher_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/teacher_micro.log
tail -n 25 gpurun_out/pytest_teacher.log | cut -c1-400; cat gpurun_out/teacher_micro.log | tail -n 12
