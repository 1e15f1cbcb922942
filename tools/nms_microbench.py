#!/usr/bin/env python
"""Device post-processing (softmax + decode + threshold compaction, per-class NMS, top-100) at large proposal counts
(BASELINE configs[4]: 256-8192 proposals x 20/80 classes), SURVEY 8(d) logits; CUDA events per entry point."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, ops  # noqa: E402
from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals  # noqa: E402


def main():
    dev = torch.device("cuda")
    B = int(os.environ.get("IMAGES", "4"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for P in [int(x) for x in os.environ.get("PROPS", "512,2048,8192").split(",")]:
        for K in (20, 80):
            gen = torch.Generator().manual_seed(99)
            lg = torch.randn(B * P, K + 1, generator=gen)
            peak = torch.rand(B * P, generator=gen) < 0.3
            cls = torch.randint(0, K, (B * P,), generator=gen)
            lg[torch.arange(B * P)[peak], cls[peak]] += 4.0
            lg[~peak, K] += 4.0
            lg, dl = lg.to(dev), (torch.randn(B * P, 4 * K, generator=gen) * 0.5).to(dev)
            pb = torch.cat([synth_proposals(P, 600, 800, torch.Generator().manual_seed(1234 + i), n_obj=8)[0] for i in range(B)], 0).to(dev)
            offs = torch.arange(0, B * P + 1, P, dtype=torch.int32, device=dev)
            hw = ops.image_hw_tensor([(600, 800)] * B, dev)
            agg = {}
            for i in range(int(os.environ.get("ITERS", "6"))):
                flush.fill_(i)
                _lib.PROFILE = {}
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                det = ops.fast_rcnn_inference_device(lg, dl, pb, offs, hw, 0.05, 0.5, 100, max_rois_per_image=P)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    agg.setdefault("total", []).append(e0.elapsed_time(e1))
                    for name, rows in _lib.PROFILE.items():
                        agg.setdefault(name, []).append(sum(a.elapsed_time(b) for a, b, _ in rows))
                _lib.PROFILE = None
            print("P=%5d K=%2d cand/img %7.0f det/img %5.1f | " % (P, K, float(det["n_candidates"].float().mean()), float(det["counts"].float().mean())) +
                  "  ".join("%s %.1f us/img" % (k.replace("b200_", ""), 1e3 * float(np.median(v)) / B) for k, v in agg.items()), flush=True)


if __name__ == "__main__":
    main()
