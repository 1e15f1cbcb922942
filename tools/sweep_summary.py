#!/usr/bin/env python
"""gpurun_out/sweep.jsonl (tools/gpu_sweep.sh) -> markdown table for profiles/."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
print("| proposals/img | classes | images/s | ms/step (4 img) | ROIAlign GB/s (of 6547 measured) | fusion GEMMs TF/s | decode + NMS us/img | res5 ms |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|")
for d in rows:
    if "failed" in d:
        print("| %s | %s | failed | | | | | |" % (d["failed"][0], d["failed"][1]))
        continue
    c, prof, st = d["config"], d["own_kernels_profile"], d["stage_ms"]
    nimg = c["images_per_gpu_per_step"]
    post = sum(prof.get(k, {}).get("ms_per_step", 0.0) for k in ("b200_softmax_decode_compact", "b200_batched_nms", "b200_gather_detections"))
    roi = d["roofline_other"] if d["roofline"]["bound"] == "tensor" else d["roofline"]
    gem = d["roofline"] if d["roofline"]["bound"] == "tensor" else d["roofline_other"]
    print("| %d | %d | %.0f | %.2f | %.0f (%.0f %%) | %.0f | %.1f | %.2f |" % (
        c["proposals_per_image"], c.get("classes", 0), d["value"], d["ms_per_step"], roi["achieved"], 100 * roi["frac"],
        gem["achieved"], 1e3 * post / nimg, st["res5_mean"]))
