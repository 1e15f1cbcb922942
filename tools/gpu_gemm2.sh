#!/bin/bash
# gemm2 tests + microbench (+ optional ncu of selected cases: NCU_CASES="conv3      N=2048,conv1 b1")
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm2.py -q -x -p no:cacheprovider --tb=short > gpurun_out/pytest_gemm2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gemm2.log
timeout 600 python tools/gemm2_microbench.py > gpurun_out/gemm2_micro.log 2>&1
echo "micro exit $?" >> gpurun_out/gemm2_micro.log
tail -n 8 gpurun_out/pytest_gemm2.log | cut -c1-250; tail -n 40 gpurun_out/gemm2_micro.log
if [ -n "$NCU_CASES" ]; then
  export CASES="$NCU_CASES" ITERS=2 TILES=256
  python tools/gemm2_microbench.py > gpurun_out/ncu_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm2_pair -s 4 -c 6 -f -o gpurun_out/prof_gemm2 python tools/gemm2_microbench.py > gpurun_out/ncu_gemm2.log 2>&1
  tail -n 5 gpurun_out/ncu_gemm2.log
fi
