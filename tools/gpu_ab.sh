#!/bin/bash
# A/B of one environment switch on the same box: usage gpu_ab.sh VAR A B [bench args]
mkdir -p gpurun_out
for v in $2 $3 $2 $3; do
  env $1=$v python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $4 > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_$v.log").read().strip().splitlines()[-1])
r = d["roofline"]
print("$1=$v", "ms %.4f" % d["ms_per_step"], "gemm b2b %.3f ms %.0f TF" % (r["ms_per_step"], r["achieved"]), "fusion fwd %.3f bwd %.3f" % (d["stage_ms"]["text_fusion_losses"], d["stage_ms"]["bwd_text_fusion"]))
PY
done
