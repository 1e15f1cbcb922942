#!/usr/bin/env python
"""SURVEY 8f-3 timing: RPN proposal selection on the C4 shapes (38x50 map x 15 anchors = 28500 anchors per image).
  train: 8 images, pre/post NMS top-k 12000 / 2000;   test: 1 image, 6000 / 1000   (NMS threshold 0.7)
Device time of the one C-ABI call (CUDA events) beside the path it replaces, run as the reference runs it: torch sort
+ gather + a per-image loop of filters and batched_nms, on the GPU with torch / torchvision ops when torchvision's CUDA
NMS is available, and on the host cores.  usage: python tools/rpn_select_microbench.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.proposal_generator import (  # noqa: E402
    find_top_rpn_proposals, find_top_rpn_proposals_device)
from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_rpn_outputs  # noqa: E402


def torch_path(props, logits, image_sizes, thr, pre, post, nms):
    """proposal_utils.py:13-118 op for op (single level)."""
    p, l = props[0], logits[0]
    N = l.shape[0]
    k = min(pre, l.shape[1])
    sl, idx = l.sort(descending=True, dim=1)
    bi = torch.arange(N, device=l.device)
    ts, ti = sl[bi, :k], idx[bi, :k]
    tb = p[bi[:, None], ti]
    out = []
    for n, (h, w) in enumerate(image_sizes):
        b, s = tb[n], ts[n]
        valid = torch.isfinite(b).all(dim=1) & torch.isfinite(s)
        if not valid.all():
            b, s = b[valid], s[valid]
        b = b.clone()
        b[:, 0::2].clamp_(min=0, max=w)
        b[:, 1::2].clamp_(min=0, max=h)
        keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
        if keep.sum().item() != len(b):
            b, s = b[keep], s[keep]
        kk = nms(b, s, thr)[:post]
        out.append((b[kk], s[kk]))
    return out


def main():
    import torchvision
    for name, N, pre, post in (("train", 8, 12000, 2000), ("test", 1, 6000, 1000)):
        gen = torch.Generator().manual_seed(9)
        props, logits = synth_rpn_outputs(N, [38 * 50 * 15], 600, 800, gen)
        sizes = [(600, 800)] * N
        dp, dl = [p.cuda() for p in props], [l.cuda() for l in logits]
        ts = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = find_top_rpn_proposals_device(dp, dl, sizes, 0.7, pre, post, 0.0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        dev_ms = sorted(ts[3:])[2]
        ws = []
        for it in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            find_top_rpn_proposals(dp, dl, sizes, 0.7, pre, post, 0.0, False)
            torch.cuda.synchronize()
            ws.append((time.perf_counter() - t0) * 1e3)
        line = "%-5s N=%d pre=%d post=%d kept=%s: b200_rpn_select_proposals %.3f ms device, %.3f ms wall as find_top_rpn_proposals" % (
            name, N, pre, post, out["counts"].tolist(), dev_ms, sorted(ws[2:])[2])
        try:
            gs = []
            for it in range(6):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                torch_path(dp, dl, sizes, 0.7, pre, post, torchvision.ops.nms)
                torch.cuda.synchronize()
                gs.append((time.perf_counter() - t0) * 1e3)
            line += "; torch/torchvision ops on the GPU %.3f ms wall" % sorted(gs[2:])[2]
        except Exception as e:  # noqa: BLE001
            line += "; torch/torchvision GPU path unavailable (%s)" % type(e).__name__
        cs = []
        for it in range(3):
            t0 = time.perf_counter()
            torch_path(props, logits, sizes, 0.7, pre, post, torchvision.ops.nms)
            cs.append((time.perf_counter() - t0) * 1e3)
        line += "; same ops on %d host threads %.1f ms" % (torch.get_num_threads(), min(cs))
        print(line, flush=True)


if __name__ == "__main__":
    main()
