#!/bin/bash
# full GPU pass: all gpu tests, smoke, default bench (train) + infer bench, ncu launch list of the default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
BENCH_DUMP_CALLS=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench_calls.log
echo "bench exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --mode infer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_infer.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log | cut -c1-600; tail -n 2 gpurun_out/bench_infer.log | cut -c1-400
