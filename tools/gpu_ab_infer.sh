#!/bin/bash
# A/B of an inference-chain switch on one box: usage gpu_ab_infer.sh [VAR] (default B200_ATTN_EPILOGUES); 0 vs 1
VAR=${1:-B200_ATTN_EPILOGUES}
mkdir -p gpurun_out
for v in 0 1 0 1; do
  env $VAR=$v python bench.py --mode infer --steps 20 --warmup 5 --no-extras --no-cpu-baseline $2 > gpurun_out/abi_$v.log 2> gpurun_out/abi_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/abi_$v.log").read().strip().splitlines()[-1])
    p = d["own_kernels_profile"]
    print("$VAR=$v", "ms %.4f" % d["ms_per_step"], "fusion+predictor %.3f" % d["stage_ms"]["text_fusion_predictor"],
          {k: round(p[k]["ms_per_step"], 3) for k in p if "attention" in k or "gemm" in k}, "launches", d["gpu_launches"])
except Exception as e:
    print("$VAR=$v failed", e); print(open("gpurun_out/abi_$v.err").read()[-800:])
PY
done
