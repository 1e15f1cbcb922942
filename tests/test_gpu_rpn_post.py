"""SURVEY 8f-3 / 8f-4 on the GPU: RPN proposal selection (find_top_rpn_proposals) and detector_postprocess through the
C ABI vs the reference's golden outputs and the oracle.  Bar: selected proposals, their order and counts bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_rpn_outputs


def T(a):
    return torch.from_numpy(np.asarray(a))


def _rpn_case(g, tag):
    N, L, pre, post = (int(v) for v in g[tag + "_meta"])
    props = [T(g["%s_props%d" % (tag, l)]) for l in range(L)]
    logits = [T(g["%s_logits%d" % (tag, l)]) for l in range(L)]
    image_sizes = [tuple(int(v) for v in s) for s in g[tag + "_image_sizes"]]
    return N, props, logits, image_sizes, float(g[tag + "_thr"]), pre, post, float(g[tag + "_min_size"])


def _device_call(props, logits, image_sizes, thr, pre, post, min_size, training=False):
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.proposal_generator import find_top_rpn_proposals
    return find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in logits], image_sizes, thr, pre, post,
                                  min_size, training)


@pytest.mark.parametrize("tag", ["c4", "c4_minsize", "fpn3"])
def test_rpn_select_golden(golden, tag):
    g = golden("rpn_select")
    N, props, logits, image_sizes, thr, pre, post, min_size = _rpn_case(g, tag)
    res = _device_call(props, logits, image_sizes, thr, pre, post, min_size)
    assert len(res) == N
    for n, r in enumerate(res):
        assert r.image_size == image_sizes[n]
        assert torch.equal(r.proposal_boxes.tensor.cpu(), T(g["%s_out_boxes%d" % (tag, n)])), (tag, n)
        assert torch.equal(r.objectness_logits.cpu(), T(g["%s_out_logits%d" % (tag, n)])), (tag, n)


def test_rpn_select_training_raises_on_nonfinite(golden):
    g = golden("rpn_select")
    N, props, logits, image_sizes, thr, pre, post, min_size = _rpn_case(g, "c4")
    with pytest.raises(FloatingPointError):
        _device_call(props, logits, image_sizes, thr, pre, post, min_size, training=True)


@pytest.mark.parametrize("N,sizes,hw,pre,post,quant", [
    (2, [38 * 50 * 15], (600, 800), 6000, 1000, 0.0),        # C4 test-time setting (RPN.PRE/POST_NMS_TOPK_TEST)
    (2, [38 * 50 * 15], (600, 800), 12000, 2000, 0.0),       # C4 train-time setting
    (1, [50 * 84 * 15], (800, 1333), 12000, 2000, 0.01),     # 800 x 1333 map, quantised logits: many exact ties
    (2, [9000, 2300, 600, 150, 40], (600, 800), 1000, 1000, 0.0),   # five levels, per-level top-k, level-batched NMS
    (1, [700], (600, 800), 6000, 1000, 0.0),                 # fewer anchors than pre_nms_topk
    (1, [20000], (600, 800), 14000, 1500, 0.0),              # a level slice beyond the cluster kernel's 12288 boxes: fallback kernel
])
def test_rpn_select_full_size_vs_oracle(N, sizes, hw, pre, post, quant):
    gen = torch.Generator().manual_seed(5 + len(sizes) + pre)
    props, logits = synth_rpn_outputs(N, sizes, hw[0], hw[1], gen, quant)
    image_sizes = [hw] * N
    ref = O.find_top_rpn_proposals(props, logits, image_sizes, 0.7, pre, post, 0.0)
    res = _device_call(props, logits, image_sizes, 0.7, pre, post, 0.0)
    for n in range(N):
        assert len(res[n].objectness_logits) == len(ref[n]["logits"]), n
        assert torch.equal(res[n].objectness_logits.cpu(), ref[n]["logits"]), n
        assert torch.equal(res[n].proposal_boxes.tensor.cpu(), ref[n]["boxes"]), n
        # size-independent properties: sorted by objectness, inside the image, pairwise IoU of survivors <= threshold
        l = res[n].objectness_logits
        assert bool((l[:-1] >= l[1:]).all())
        b = res[n].proposal_boxes.tensor
        assert float(b.min()) >= 0 and float(b[:, 0::2].max()) <= hw[1] and float(b[:, 1::2].max()) <= hw[0]
    if len(sizes) == 1:
        from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, pairwise_iou
        b = Boxes(res[0].proposal_boxes.tensor[:400])
        iou = pairwise_iou(b, b)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.7


def test_rpn_select_all_equal_logits_and_empty():
    """Every logit equal: the tie rule (lower anchor index first) decides the whole selection; and a level of zero anchors."""
    gen = torch.Generator().manual_seed(11)
    props, logits = synth_rpn_outputs(1, [5000], 600, 800, gen)
    logits = [torch.zeros_like(logits[0])]
    logits[0][0, ::7] = -0.0                                  # -0 == +0 for the sort
    ref = O.find_top_rpn_proposals(props, logits, [(600, 800)], 0.7, 1000, 300, 0.0)
    res = _device_call(props, logits, [(600, 800)], 0.7, 1000, 300, 0.0)
    assert torch.equal(res[0].proposal_boxes.tensor.cpu(), ref[0]["boxes"])
    res = _device_call([torch.zeros(2, 0, 4)], [torch.zeros(2, 0)], [(600, 800)] * 2, 0.7, 1000, 300, 0.0)
    assert all(len(r.objectness_logits) == 0 for r in res)


def test_detector_postprocess_vs_oracle():
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.postprocessing import (detector_postprocess,
                                                                                               detector_postprocess_batch)
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances
    gen = torch.Generator().manual_seed(3)
    sizes = [((600, 800), (375, 500)), ((608, 913), (333, 500)), ((480, 672), (960, 1344))]
    N, topk = len(sizes), 100
    boxes = torch.zeros(N, topk, 4)
    scores = torch.zeros(N, topk)
    classes = torch.full((N, topk), -1, dtype=torch.int64)
    roi = torch.full((N, topk), -1, dtype=torch.int64)
    counts = torch.tensor([100, 37, 0], dtype=torch.int32)
    for n, ((h, w), _) in enumerate(sizes):
        c = int(counts[n])
        b = torch.rand(c, 4, generator=gen) * torch.tensor([w, h, w, h])
        b[:, 2:] = torch.minimum(b[:, :2] + torch.rand(c, 2, generator=gen) * 200, torch.tensor([float(w), float(h)]))
        if c > 8:
            b[3, 2] = b[3, 0]                                  # empty before scaling
        boxes[n, :c] = b
        scores[n, :c] = torch.rand(c, generator=gen).sort(descending=True).values
        classes[n, :c] = torch.randint(0, 20, (c,), generator=gen)
        roi[n, :c] = torch.randperm(512, generator=gen)[:c]
    det = dict(boxes=boxes.cuda(), scores=scores.cuda(), classes=classes.cuda(), roi_inds=roi.cuda(), counts=counts.cuda())
    detector_postprocess_batch(det, [s[0] for s in sizes], [s[1] for s in sizes])
    for n, ((h, w), (oh, ow)) in enumerate(sizes):
        c = int(counts[n])
        rb, keep = O.detector_postprocess(boxes[n, :c], (h, w), oh, ow)
        k = int(det["counts"][n])
        assert k == int(keep.sum())
        assert torch.equal(det["boxes"][n, :k].cpu(), rb)
        assert torch.equal(det["scores"][n, :k].cpu(), scores[n, :c][keep])
        assert torch.equal(det["classes"][n, :k].cpu(), classes[n, :c][keep])
        assert torch.equal(det["roi_inds"][n, :k].cpu(), roi[n, :c][keep])
        assert bool((det["classes"][n, k:] == -1).all())
    # Instances form (detectron2's signature)
    (h, w), (oh, ow) = sizes[0]
    inst = Instances((h, w))
    inst.pred_boxes = Boxes(boxes[0].cuda())
    inst.scores = scores[0].cuda()
    inst.pred_classes = classes[0].cuda()
    out = detector_postprocess(inst, oh, ow)
    rb, keep = O.detector_postprocess(boxes[0], (h, w), oh, ow)
    assert out.image_size == (oh, ow)
    assert torch.equal(out.pred_boxes.tensor.cpu(), rb)
    assert torch.equal(out.scores.cpu(), scores[0][keep]) and torch.equal(out.pred_classes.cpu(), classes[0][keep])


def test_rpn_select_adversarial_grid_cases():
    """Integer-grid anchors (duplicates, IoU exactly at the threshold, boxes that clip to nothing or fall under min_size)
    and logits in 0.25 steps (long exact-tie runs across the pre_nms_topk boundary): selection and order equal the
    oracle's stable-sort restatement."""
    gen = torch.Generator().manual_seed(31)
    for case in range(12):
        N = 1 + case % 3
        sizes = [[900], [400, 100], [64, 32, 16]][case % 3]
        A = sum(sizes)
        xy = torch.randint(-4, 40, (N, A, 2), generator=gen).float() * 4
        wh = torch.randint(0, 10, (N, A, 2), generator=gen).float() * 4
        boxes = torch.cat([xy, xy + wh], 2)
        lg = torch.randint(-8, 9, (N, A), generator=gen).float() / 4
        props = list(boxes.split(sizes, dim=1))
        logits = list(lg.split(sizes, dim=1))
        pre, post = [50, 300, 1000][case % 3], [20, 1000, 100][(case // 3) % 3]
        thr, ms = [0.5, 0.7, 0.25][case % 3], [0.0, 4.0][case % 2]
        image_sizes = [(120, 150)] * N
        ref = O.find_top_rpn_proposals(props, logits, image_sizes, thr, pre, post, ms)
        res = _device_call(props, logits, image_sizes, thr, pre, post, ms)
        for n in range(N):
            assert torch.equal(res[n].objectness_logits.cpu(), ref[n]["logits"]), (case, n)
            assert torch.equal(res[n].proposal_boxes.tensor.cpu(), ref[n]["boxes"]), (case, n)
