#!/bin/bash
# default bench (train) + infer bench, then a full ncu capture of the RPN selection kernels in the microbench
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --mode infer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_infer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rpn_topk_filter|nms_presorted_cluster' -c 4 -f -o gpurun_out/prof_rpn python tools/rpn_select_microbench.py > gpurun_out/ncu_rpn.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_rpn.log
tail -n 2 gpurun_out/bench.log | cut -c1-300; tail -n 3 gpurun_out/bench.err; tail -n 2 gpurun_out/bench_infer.log | cut -c1-300; tail -n 3 gpurun_out/ncu_rpn.log
