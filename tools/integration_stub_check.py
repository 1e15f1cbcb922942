#!/usr/bin/env python
"""Executes the ctypes stub of INTEGRATION.md section B as written and checks it against the package's own wrappers.
usage: python tools/integration_stub_check.py"""
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200 import ops  # noqa: E402
from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals, synth_rpn_outputs  # noqa: E402

md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
block = next(b for b in re.findall(r"```python\n(.*?)```", md, re.S) if "ctypes.CDLL" in b)
ns = {}
exec(block, ns)
gen = torch.Generator().manual_seed(0)
feat = torch.relu(torch.randn(2, 64, 38, 50, generator=gen)).cuda()
b = [synth_proposals(40, 600, 800, gen)[0].cuda() for _ in range(2)]
rois, _ = ops.boxes_to_rois(b)
a = ns["roi_align"](feat, rois)
w = ops.roi_align(feat, rois, (7, 7), 1 / 16, 0, True)
assert torch.equal(a, w), "roi_align stub"
boxes = synth_proposals(500, 600, 800, gen)[0].cuda()
scores, idxs = torch.rand(500, generator=gen).cuda(), torch.randint(0, 20, (500,), generator=gen).cuda()
assert torch.equal(ns["batched_nms"](boxes, scores, idxs, 0.5), ops.batched_nms(boxes, scores, idxs, 0.5)), "batched_nms stub"
props, logits = synth_rpn_outputs(2, [3000], 600, 800, gen)
hw = ops.image_hw_tensor([(600, 800)] * 2, "cuda")
bx, lg, cnt, bad = ns["rpn_select"](props[0].cuda().contiguous(), logits[0].cuda().contiguous(), hw, 0.7, 1000, 300)
ref = ops.rpn_select_proposals(props[0].cuda(), logits[0].cuda(), [3000], hw, 0.7, 1000, 300)
assert torch.equal(bx, ref["boxes"]) and torch.equal(lg, ref["logits"]) and torch.equal(cnt, ref["counts"]), "rpn_select stub"
x = (torch.randn(1000, 512, generator=gen) * 0.5).to(torch.bfloat16).cuda()
wgt = (torch.randn(384, 512, generator=gen) * 0.05).to(torch.bfloat16).cuda()
bias = torch.randn(384, generator=gen).cuda()
res = torch.randn(1000, 384, generator=gen).to(torch.bfloat16).cuda()
y = ns["linear_relu"](x, wgt, bias, res)
assert torch.equal(y, ops.gemm2(x, wgt, bias=bias, residual=res, relu=True)), "gemm2 stub"
ref_y = torch.relu(x.double() @ wgt.double().t() + bias.double() + res.double())
torch.testing.assert_close(y.double(), ref_y, rtol=1e-2, atol=1e-2)
import ctypes  # noqa: E402
from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib  # noqa: E402
assert ctypes.sizeof(ns["Gemm2Desc"]) == ctypes.sizeof(_lib.Gemm2Desc), "Gemm2Desc layout"
print("INTEGRATION.md stubs ok:", a.shape, int(cnt.sum()), tuple(y.shape))
