#!/bin/bash
# full ncu capture of one whole step's worth of the three headline kernels (GEMM chain, ROIAlign forward, ROIAlign
# backward gather) inside a short bench run; the plain run goes first
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05_kernel|roi_align_fwd_slice_kernel|roi_bwd_csr_gather_kernel" -s 84 -c 28 -f -o gpurun_out/prof_top $BENCH > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
