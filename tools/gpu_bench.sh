#!/bin/bash
# default bench line (+ extras), stderr kept
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 $BENCH_ARGS > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"
tail -n 15 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["cuda_graph"])
    print("stage", d["stage_ms"])
    r = d["roofline"]; print("roofline", r["achieved"], r["frac"], r["ms_per_step"], "in_step", r["in_step"]["achieved"], r["res5_convolutions"])
    print("roi", d["roofline_other"]["achieved"], d["roofline_other"]["standalone_op"]["fwd_frac"], d["roofline_other"]["standalone_op"]["bwd_frac"])
    for k, v in d.get("extra", {}).items(): print(k, {kk: vv for kk, vv in v.items() if kk not in ("config",)})
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("parse failed", e)
PY
