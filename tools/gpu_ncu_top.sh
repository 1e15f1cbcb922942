#!/bin/bash
# full ncu capture of one whole step's worth of the headline kernels (tensor-core GEMMs, ROIAlign forward, ROIAlign backward gather,
# SGD) inside a short eager bench run; the plain run goes first.  48 matching launches per fine-tune step: skip two steps, take one.
# The raw page is exported on the box (the report itself only travels back when it fits gpurun_out's 64 MiB).
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"gemm2_pair_kernel|gemm_bf16_tcgen05_kernel|roi_align_fwd_slice_kernel|roi_bwd_csr_gather_kernel|sgd_momentum_kernel" -s ${NCU_SKIP:-96} -c ${NCU_COUNT:-48} -f -o /tmp/prof_top $BENCH > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i /tmp/prof_top.ncu-rep --page raw --csv > gpurun_out/prof_top_raw.csv 2> gpurun_out/prof_top_raw.err
ls -la /tmp/prof_top.ncu-rep gpurun_out/prof_top_raw.csv
[ $(stat -c %s /tmp/prof_top.ncu-rep) -lt 40000000 ] && cp /tmp/prof_top.ncu-rep gpurun_out/prof_top.ncu-rep
du -sh gpurun_out
