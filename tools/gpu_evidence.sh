#!/bin/bash
# round-2 side evidence: gradient parity table, NMS / post-processing microbench at 8192 x 80, gemm2 microbench
mkdir -p gpurun_out
timeout 600 python tools/grad_parity_table.py > gpurun_out/grad_parity_table.md 2> gpurun_out/grad_parity_table.err; echo "grad table exit $?"; tail -n 3 gpurun_out/grad_parity_table.err
timeout 600 python tools/nms_microbench.py > gpurun_out/nms_microbench.txt 2>&1; echo "nms exit $?"; tail -n 12 gpurun_out/nms_microbench.txt
timeout 600 python tools/gemm2_microbench.py > gpurun_out/gemm2_micro.log 2>&1; echo "gemm2 exit $?"
