#!/bin/bash
# gemm2 tests + the old GEMM tests (shared header refactor) + microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm2.py tests/test_gpu_fusion.py -q -x -p no:cacheprovider --tb=short > gpurun_out/pytest_gemm2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gemm2.log
timeout 600 python tools/gemm2_microbench.py > gpurun_out/gemm2_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/gemm2_micro.log
tail -n 30 gpurun_out/pytest_gemm2.log | cut -c1-250; tail -n 40 gpurun_out/gemm2_micro.log
