#!/bin/bash
# new-kernel pass: RPN selection / detector_postprocess tests, the whole gpu suite, the RPN microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rpn_post.py -q -p no:cacheprovider --tb=short -x > gpurun_out/pytest_rpn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_rpn.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/rpn_select_microbench.py > gpurun_out/rpn_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/rpn_micro.log
tail -n 15 gpurun_out/pytest_rpn.log; tail -n 4 gpurun_out/pytest_gpu.log; cat gpurun_out/rpn_micro.log
