#!/bin/bash
# A/B of the inference chain: attention probabilities / gate operands as GEMM epilogues (1) vs the separate attention kernel (0)
mkdir -p gpurun_out
for v in 0 1 0 1; do
  env B200_ATTN_EPILOGUES=$v python bench.py --mode infer --steps 20 --warmup 5 --no-extras --no-cpu-baseline $1 > gpurun_out/abi_$v.log 2> gpurun_out/abi_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/abi_$v.log").read().strip().splitlines()[-1])
p = d["own_kernels_profile"]
print("B200_ATTN_EPILOGUES=$v", "ms %.4f" % d["ms_per_step"], "fusion+predictor %.3f" % d["stage_ms"]["text_fusion_predictor"],
      {k: round(p[k]["ms_per_step"], 3) for k in p if "attention" in k or "gemm" in k}, "launches", d["gpu_launches"])
PY
done
