#!/bin/bash
# res5 on own kernels: parity tests, then the default bench line with both implementations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_res5.py tests/test_gpu_gemm2.py -q -x -p no:cacheprovider --tb=short > gpurun_out/pytest_res5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_res5.log
tail -n 30 gpurun_out/pytest_res5.log | cut -c1-300
for impl in tcgen05 cudnn; do
  B200_RES5_IMPL=$impl timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_res5_$impl.log 2> gpurun_out/bench_res5_$impl.err
  echo "bench $impl exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_res5_$impl.log").read().strip().splitlines()[-1])
    print("$impl", d["value"], d["ms_per_step"], d["stage_ms"], d["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
  tail -n 3 gpurun_out/bench_res5_$impl.err
done
