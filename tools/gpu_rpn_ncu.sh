#!/bin/bash
# per-kernel times of the RPN selection (launch list of the microbench)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rpn_post.py -q -p no:cacheprovider --tb=short -x > gpurun_out/pytest_rpn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_rpn.log
timeout 300 python tools/rpn_select_microbench.py > gpurun_out/rpn_micro.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rpn_|nms_' -c 60 --csv --log-file gpurun_out/rpn_launches.csv python tools/rpn_select_microbench.py > gpurun_out/rpn_ncu.log 2>&1
tail -n 3 gpurun_out/pytest_rpn.log; cat gpurun_out/rpn_micro.log; tail -n 30 gpurun_out/rpn_launches.csv | cut -d, -f5,12- 
