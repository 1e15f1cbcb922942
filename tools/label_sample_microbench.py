#!/usr/bin/env python
"""S1 timing on the training shape: 8 images x 2000 RPN proposals (+ ground truth appended), 8 objects per image.
Kernel path (one launch + one small D2H read) vs the torch-op path it replaces (per-image pairwise_iou / Matcher /
subsample_labels, as the reference runs it), through ROIHeads.label_and_sample_proposals."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling  # noqa: E402
from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec  # noqa: E402
from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals  # noqa: E402


def main():
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME, cfg.MODEL.ROI_HEADS.NUM_CLASSES = "SematicRes5ROIHeads", 20
    cfg.MODEL.ADDITION.NAME = "clip"
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 4, 1
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=16, stride=16)}).cuda().train()
    props, targets = [], []
    for i in range(8):
        gen = torch.Generator().manual_seed(100 + i)
        b, objs = synth_proposals(2000, 600, 800, gen, n_obj=8)
        p = Instances((600, 800))
        p.proposal_boxes = Boxes(b.cuda())
        p.objectness_logits = torch.zeros(2000, device="cuda")
        t = Instances((600, 800))
        t.gt_boxes = Boxes(objs.cuda())
        t.gt_classes = torch.randint(0, 20, (8,), generator=gen).cuda()
        props.append(p)
        targets.append(t)
    for name, fn in (("kernel", lambda: m.label_and_sample_proposals(props, targets)), ("torch ops", lambda: _torch_path(m, props, targets))):
        ts = []
        for it in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        print("%-10s %.3f ms per call (8 images x 2008 proposals, wall clock incl. host work; median of last 5: %.3f)" %
              (name, min(ts), sorted(ts[3:])[2]))
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    from fewshotobjectdetection_imporove_via_text_feature_b200.layers import add_ground_truth_to_proposals
    pp = add_ground_truth_to_proposals([t.gt_boxes for t in targets], props)
    args = ([p.proposal_boxes.tensor for p in pp], [t.gt_boxes.tensor for t in targets], [t.gt_classes for t in targets], 20)
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.label_and_sample_proposals(*args, seed=it)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("device time of ops.label_and_sample_proposals (concatenation + kernel): %.3f ms" % sorted(ts[3:])[2])


def _torch_path(m, proposals, targets):
    from fewshotobjectdetection_imporove_via_text_feature_b200.layers import add_ground_truth_to_proposals
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import pairwise_iou
    proposals = add_ground_truth_to_proposals([t.gt_boxes for t in targets], proposals)
    out = []
    for p, t in zip(proposals, targets):
        idxs, labels = m.proposal_matcher(pairwise_iou(t.gt_boxes, p.proposal_boxes))
        sampled, gt_classes = m._sample_proposals(idxs, labels, t.gt_classes)
        q = p[sampled]
        q.gt_classes = gt_classes
        q.gt_boxes = Boxes(t.gt_boxes.tensor[idxs[sampled]])
        n_bg = (gt_classes == m.num_classes).sum().item()
        out.append(q)
    return out


if __name__ == "__main__":
    main()
