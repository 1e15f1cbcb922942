#!/bin/bash
# A/B of one environment switch on the same box: usage gpu_ab.sh VAR A B [bench args]
mkdir -p gpurun_out
for v in $2 $3 $2 $3 $2 $3; do
  env $1=$v python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $4 > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$v.log").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$1=$v", "ms %.4f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "gemm b2b %.3f ms %.0f TF" % (r["ms_per_step"], r["achieved"]))
except Exception as e:
    print("$1=$v failed", e); print(open("gpurun_out/ab_$v.err").read()[-1200:])
PY
done
