#!/usr/bin/env python
"""ROIAlign forward microbench on the bench shape (B images x P proposals, 1024 x 38 x 50 bf16 channels-last):
every implementation x bin_step, CUDA-event timed with an L2 flush between launches."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, ops  # noqa: E402
from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--props", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--impls", default="0,2")
    ap.add_argument("--steps", default="1,2")
    ap.add_argument("--variants", default="0,1,2,3,4", help="backward variants to time (see bench_bwd)")
    ap.add_argument("--tile-variants", default="0", help="variants of the tile gather to time (roi_bwd_tile_variant: 0 default = heaviest tile first, 1 / 2 pipelining, 3 map order)")
    ap.add_argument("--bwd", action="store_true", help="time the backward (b200_roi_align_bwd entry point) instead")
    a = ap.parse_args()
    B, P, C, H, W = a.images, a.props, 1024, 38, 50
    dev = torch.device("cuda")
    feat = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(0))).to(torch.bfloat16)
    feat = feat.to(dev).contiguous(memory_format=torch.channels_last)
    boxes = [synth_proposals(P, 600, 800, torch.Generator().manual_seed(1234 + i), n_obj=8)[0].to(dev) for i in range(B)]
    rois, offs = ops.boxes_to_rois(boxes)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if a.bwd:
        return bench_bwd(a, feat, rois, offs, flush)
    outs = {}
    for step in [int(s) for s in a.steps.split(",")]:
        nb = -(-7 // step)
        nbytes = B * C * H * W * 2 + B * P * 20 + B * P * C * nb * nb * 2
        for impl in [int(s) for s in a.impls.split(",")]:
            _lib.set_option("roi_align_bf16_impl", impl)
            ts = []
            for i in range(a.iters + 3):
                flush.fill_(i & 255)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = ops.roi_align(feat, rois, 7, 1 / 16, 0, True, channels_last_out=True, roi_batch_offsets=offs,
                                    bin_step=step)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            outs[(step, impl)] = out.float()
            print("bin_step=%d impl=%d  %.4f ms  %.1f GB/s (algorithmic %.1f MB)  min %.4f ms" %
                  (step, impl, ms, nbytes / ms / 1e6, nbytes / 1e6, min(ts)), flush=True)
        ks = [k for k in outs if k[0] == step]
        for k in ks[1:]:
            d = (outs[k] - outs[ks[0]]).abs().max().item()
            print("   max |impl %d - impl %d| = %.4g (ref max %.3g)" % (k[1], ks[0][1], d, outs[ks[0]].abs().max().item()))
    _lib.set_option("roi_align_bf16_impl", 2)


def bench_bwd(a, feat, rois, offs, flush):
    B, C, H, W = feat.shape
    P = a.props
    for step in [int(s) for s in a.steps.split(",")]:
        nb = -(-7 // step)
        nbytes = B * C * H * W * 2 + B * P * 20 + B * P * C * nb * nb * 2
        g = torch.randn(B * P, C, nb, nb, device=feat.device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        res = {}
        # 0: fp32-table kernel, 1: CSR lists built inside the call, 2: CSR lists planned ahead,
        # 3: pixel-tile tensor-core gather with the plan built inside the call, 4: tile plan built ahead
        variants = [int(v) for v in a.variants.split(",")]
        tvs = [int(v) for v in a.tile_variants.split(",")]
        for impl, tv in [(i, 0) for i in variants if i != 4] + [(4, v) for v in tvs if 4 in variants]:
            _lib.set_option("roi_align_bwd_impl", {0: 0, 1: 1, 2: 1, 3: 2, 4: 2}[impl])
            _lib.set_option("roi_bwd_tile_variant", tv)
            ops.PLAN_AHEAD[0] = impl in (2, 4)
            x = feat.clone().requires_grad_(True)
            out = ops.roi_align(x, rois, 7, 1 / 16, 0, True, channels_last_out=True, roi_batch_offsets=offs, bin_step=step)
            ts = []
            for i in range(a.iters + 3):
                flush.fill_(i & 255)
                x.grad = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out.backward(g, retain_graph=True)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            if impl == 4 and tv != tvs[0]:
                print("   tile variant %d == variant %d: %s" % (tv, tvs[0], torch.equal(res[4], x.grad.float())))
            else:
                res[impl] = x.grad.float()
            print("bwd bin_step=%d impl=%d tv=%d  %.4f ms  %.1f GB/s (algorithmic %.1f MB)  min %.4f ms" %
                  (step, impl, tv, ms, nbytes / ms / 1e6, nbytes / 1e6, min(ts)), flush=True)
        _lib.set_option("roi_bwd_tile_variant", 0)
        if len(variants) < 5:
            continue
        d = (res[1] - res[0]).norm() / res[0].norm()
        d3 = (res[3] - res[0]).norm() / res[0].norm()
        print("   rel |impl 1 - impl 0| = %.3g, impl 2 == impl 1: %s; rel |impl 3 - impl 0| = %.3g, impl 4 == impl 3: %s" %
              (d.item(), torch.equal(res[1], res[2]), d3.item(), torch.equal(res[3], res[4])))
    _lib.set_option("roi_align_bwd_impl", _lib.ROI_BWD_IMPL_DEFAULT)
    ops.PLAN_AHEAD[0] = True


if __name__ == "__main__":
    main()
