"""Decode / threshold compaction / per-class NMS kernels vs the oracle and the reference golden vectors.
Bar (north_star): NMS keep indices and compaction counts bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from oracle.gen_golden import synth_proposals


def T(a):
    return torch.from_numpy(np.asarray(a))


def _run(logits_or_probs, deltas, props, hw, is_prob, topk=100):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    R = logits_or_probs.shape[0]
    offs = torch.tensor([0, R], dtype=torch.int32, device="cuda")
    ihw = torch.tensor([[float(hw[0]), float(hw[1])]], device="cuda")
    return ops.fast_rcnn_inference_device(logits_or_probs.cuda(), deltas.cuda(), props.cuda(), offs, ihw, 0.05, 0.5, topk,
                                          input_is_prob=is_prob, want_probs=True)


@pytest.mark.parametrize("tag", ["voc", "coco", "ties", "empty"])
def test_golden_fused_path(golden, tag):
    g = golden("fast_rcnn_inference")
    hw = g[tag + "_hw"]
    # (1) probabilities in: the candidate list must be exactly the reference's nonzero() list
    out = _run(T(g[tag + "_probs"]), T(g[tag + "_deltas"]), T(g[tag + "_props"]), hw, True)
    n = int(out["n_candidates"][0])
    assert n == int(g[tag + "_ncand"])
    ref_idx = (T(g[tag + "_probs"])[:, :-1] > 0.05).nonzero()
    c = out["cand"]
    assert torch.equal(c["cand_roi"][:n].cpu().long(), ref_idx[:, 0])
    assert torch.equal(c["cand_cls"][:n].cpu().long(), ref_idx[:, 1])
    assert torch.equal(c["cand_scores"][:n].cpu(), T(g[tag + "_probs"])[:, :-1][T(g[tag + "_probs"])[:, :-1] > 0.05])
    # decoded + clipped candidate boxes vs the reference's Box2BoxTransform (expf ulp differences allowed)
    refb = T(g[tag + "_pred_boxes_all"]).reshape(len(g[tag + "_props"]), -1, 4).clone()
    refb[..., 0::2] = refb[..., 0::2].clamp(0, float(hw[1]))
    refb[..., 1::2] = refb[..., 1::2].clamp(0, float(hw[0]))
    if n:
        torch.testing.assert_close(c["cand_boxes"][:n].cpu(), refb[ref_idx[:, 0], ref_idx[:, 1]], rtol=1e-5, atol=2e-3)
    # (2) NMS is bit-exact on the kernel's own candidates
    keep_ref = O.batched_nms(c["cand_boxes"][:n].cpu(), c["cand_scores"][:n].cpu(), c["cand_cls"][:n].cpu().long(), 0.5)[:100]
    k = int(out["counts"][0])
    assert k == len(keep_ref)
    assert torch.equal(out["keep"][0, :k].cpu().long(), keep_ref)
    # (3) and the final detections equal the reference's (same classes / rois; scores exact; boxes to ulp)
    assert k == len(g[tag + "_scores"])
    assert torch.equal(out["classes"][0, :k].cpu(), T(g[tag + "_classes"]))
    assert torch.equal(out["roi_inds"][0, :k].cpu(), T(g[tag + "_roi_inds"]))
    assert torch.equal(out["scores"][0, :k].cpu(), T(g[tag + "_scores"]))
    if k:
        torch.testing.assert_close(out["boxes"][0, :k].cpu(), T(g[tag + "_boxes"]), rtol=1e-5, atol=2e-3)
    # (4) logits in: softmax within 1e-6 of torch's
    out2 = _run(T(g[tag + "_logits"]), T(g[tag + "_deltas"]), T(g[tag + "_props"]), hw, False)
    torch.testing.assert_close(out2["probs"].cpu(), T(g[tag + "_probs"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["voc", "coco", "ties", "empty"])
def test_golden_single_image_entry(golden, tag):
    """Reference-signature entry on the reference's own decoded boxes / probabilities: bit-exact everything."""
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling import roi_heads as RH
    g = golden("fast_rcnn_inference")
    res, roi = RH.fast_rcnn_inference_single_image(T(g[tag + "_pred_boxes_all"]).cuda(), T(g[tag + "_probs"]).cuda(),
                                                   tuple(int(v) for v in g[tag + "_hw"]), 0.05, 0.5, 100)
    assert torch.equal(res.pred_classes.cpu(), T(g[tag + "_classes"]))
    assert torch.equal(roi.cpu(), T(g[tag + "_roi_inds"]))
    assert torch.equal(res.scores.cpu(), T(g[tag + "_scores"]))
    assert torch.equal(res.pred_boxes.tensor.cpu(), T(g[tag + "_boxes"]))


@pytest.mark.parametrize("n,ncls,seed", [(1, 1, 0), (2, 1, 1), (257, 5, 2), (3000, 20, 3), (9000, 80, 4), (6000, 1, 5),
                                         (4096, 1, 6), (4097, 1, 7), (8192, 1, 8), (8193, 1, 9), (20000, 3, 10)])
def test_batched_nms_vs_oracle(n, ncls, seed):
    """Class slices of every size class of the kernel: <= 4096 boxes (256-thread CTA, shared memory), 4097-8192 (1024-thread
    CTA, 197 KB shared memory), above (global-memory path)."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(seed)
    boxes, _ = synth_proposals(max(n, 8), 600, 800, gen, n_obj=6)
    boxes = boxes[:n].contiguous()
    scores = torch.rand(n, generator=gen)
    if n > 40:
        scores[5:25] = scores[0]                 # ties -> ascending-index order
        boxes[30:40] = boxes[20:30]              # identical boxes
    idxs = torch.randint(0, ncls, (n,), generator=gen)
    ref = O.batched_nms(boxes, scores, idxs, 0.5)
    keep = ops.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5)
    assert torch.equal(keep.cpu(), ref)


@pytest.mark.parametrize("n,ncls,seed", [(39999, 80, 11), (40000, 80, 12), (45000, 20, 13)])
def test_batched_nms_around_the_40000_box_rule(n, ncls, seed):
    """detectron2 v0.3's batched_nms changes algorithm at 40 000 boxes (fast_rcnn.py:125 -> layers/nms.py): coordinate
    trick below, per-class NMS on the un-offset boxes from there on.  Keep indices bit-exact on both sides of the rule,
    at 80 classes (BASELINE configs[2] / [4])."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(seed)
    boxes, _ = synth_proposals(n, 600, 800, gen, n_obj=12)
    scores = torch.rand(n, generator=gen)
    boxes[300:340] = boxes[200:240]              # identical boxes
    idxs = torch.randint(0, ncls, (n,), generator=gen)
    ref = O.batched_nms_detectron2(boxes, scores, idxs, 0.5)
    keep = ops.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5)
    assert torch.equal(keep.cpu(), ref)


def test_one_class_takes_every_proposal_large_image():
    """BASELINE configs[4] with an untrained classifier: one class passes the threshold on (nearly) all 8192 proposals of an
    image, i.e. a single 8192-box class slice per image (the 1024-thread / 197 KB instantiation of the per-class kernel).
    NMS bit-exact against the oracle on the kernel's own candidates, same result with and without the ROI-count hint."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(21)
    P, K, N = 8192, 20, 2
    boxes = torch.cat([synth_proposals(P, 600, 800, gen, n_obj=8)[0] for _ in range(N)])
    logits = torch.randn(N * P, K + 1, generator=gen) * 0.3
    logits[:, 3] += 5.0                                   # class 3 wins everywhere
    logits[::7, 11] += 4.5                                # a second class on every 7th proposal
    deltas = torch.randn(N * P, 4 * K, generator=gen) * 0.3
    offs = torch.arange(0, N * P + 1, P, dtype=torch.int32)
    hw = torch.tensor([[600.0, 800.0]] * N)
    outs = []
    for hint in (P, None):
        outs.append(ops.fast_rcnn_inference_device(logits.cuda(), deltas.cuda(), boxes.cuda(), offs.cuda(), hw.cuda(), 0.05, 0.5, 100,
                                                   want_probs=True, max_rois_per_image=hint))
    for k in ("boxes", "scores", "classes", "roi_inds", "counts", "keep"):
        assert torch.equal(outs[0][k], outs[1][k]), k
    out, c = outs[0], outs[0]["cand"]
    for i in range(N):
        n = int(out["n_candidates"][i])
        s0 = i * P * K
        cc = c["cand_cls"][s0:s0 + n].cpu().long()
        assert int((cc == 3).sum()) == P                   # the whole image in one class slice
        ref = O.batched_nms(c["cand_boxes"][s0:s0 + n].cpu(), c["cand_scores"][s0:s0 + n].cpu(), cc, 0.5)[:100]
        k = int(out["counts"][i])
        assert k == len(ref) == 100 and torch.equal(out["keep"][i, :k].cpu().long(), ref)


def test_batched_nms_empty_and_multi_segment():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    assert ops.batched_nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), 0.5).numel() == 0
    gen = torch.Generator().manual_seed(3)
    segs, refs = [], []
    caps = [700, 0, 1300, 64]
    for n in caps:
        b, _ = synth_proposals(max(n, 8), 480, 640, gen)
        b = b[:n]
        s, c = torch.rand(n, generator=gen), torch.randint(0, 20, (n,), generator=gen)
        segs.append((b, s, c))
        refs.append(O.batched_nms(b, s, c, 0.5)[:100])
    pad = 50                                       # sparse segments: capacity > count
    offs, boxes, scores, cls = [0], [], [], []
    for b, s, c in segs:
        boxes += [b, torch.zeros(pad, 4)]
        scores += [s, torch.zeros(pad)]
        cls += [c, torch.zeros(pad, dtype=torch.int64)]
        offs.append(offs[-1] + len(b) + pad)
    keep, kc = ops.batched_nms_segments(torch.cat(boxes).cuda(), torch.cat(scores).cuda(), torch.cat(cls).to(torch.int32).cuda(),
                                        torch.tensor(offs[:-1], dtype=torch.int32).cuda(),
                                        torch.tensor(caps, dtype=torch.int32).cuda(), 20, 0.5, 100)
    for i, r in enumerate(refs):
        assert int(kc[i]) == len(r)
        assert torch.equal(keep[i, :len(r)].cpu().long(), r)


def test_multi_image_full_size():
    """BASELINE shape: 4 images x 512 proposals, K=20.  Per-image results equal the single-image oracle on the
    kernel's own probabilities/candidates; idempotence: NMS of the kept set keeps everything."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(0)
    N, R, K = 4, 512, 20
    props = torch.cat([synth_proposals(R, 600, 800, gen)[0] for _ in range(N)])
    logits = torch.randn(N * R, K + 1, generator=gen)
    logits[:, K] += 3.0
    pk = torch.randint(0, N * R, (N * R // 3,), generator=gen)
    logits[pk, torch.randint(0, K, (len(pk),), generator=gen)] += 7.0
    deltas = torch.randn(N * R, 4 * K, generator=gen) * 0.5
    offs = torch.arange(0, N + 1, dtype=torch.int32) * R
    hw = torch.tensor([[600.0, 800.0]] * N)
    out = ops.fast_rcnn_inference_device(logits.cuda(), deltas.cuda(), props.cuda(), offs.cuda(), hw.cuda(), 0.05, 0.5, 100,
                                         want_probs=True)
    probs = out["probs"].cpu()
    c = out["cand"]
    for i in range(N):
        n = int(out["n_candidates"][i])
        assert n == int((probs[i * R:(i + 1) * R, :K] > 0.05).sum())
        s0 = i * R * K
        cb, cs, cc = c["cand_boxes"][s0:s0 + n].cpu(), c["cand_scores"][s0:s0 + n].cpu(), c["cand_cls"][s0:s0 + n].cpu().long()
        ref = O.batched_nms(cb, cs, cc, 0.5)[:100]
        k = int(out["counts"][i])
        assert k == len(ref) and torch.equal(out["keep"][i, :k].cpu().long(), ref)
        assert bool((out["scores"][i, :k - 1] >= out["scores"][i, 1:k]).all())          # sorted
        again = ops.batched_nms(out["boxes"][i, :k], out["scores"][i, :k], out["classes"][i, :k], 0.5)
        assert torch.equal(again.cpu(), torch.arange(k))                                 # idempotent


def test_adversarial_grid_cases_bit_exact():
    """Integer-grid boxes (IoU exactly 0.5 after the class offset, duplicates, boxes that clip to nothing), probabilities in
    0.05 steps (long runs of exact ties, values exactly at the score threshold): candidates, keep order and counts
    must equal the oracle's — itself pinned on the reference's own function for this family of cases
    (tests/test_oracle_pinning.py)."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(2025)
    for case in range(30):
        R = int(torch.randint(1, 200, (1,), generator=gen))
        K = [3, 20, 1, 80][case % 4]
        xy = torch.randint(-3, 30, (R, 2), generator=gen).float() * 4
        wh = torch.randint(0, 8, (R, 2), generator=gen).float() * 4 + (4 if case % 5 else 0)
        props = torch.cat([xy, xy + wh], 1)
        probs = torch.randint(0, 21, (R, K + 1), generator=gen).float() / 20
        hw = (100, 120)
        topk = [100, 7][case % 2]
        deltas = torch.zeros(R, 4 * K)
        out = _run(probs, deltas, props, hw, True, topk)
        # zero deltas decode to the proposal itself only where that is exact in fp32: take the kernel's own decoded boxes
        n = int(out["n_candidates"][0])
        c = out["cand"]
        ref_idx = (probs[:, :-1] > 0.05).nonzero()
        assert n == len(ref_idx), case
        assert torch.equal(c["cand_roi"][:n].cpu().long(), ref_idx[:, 0]) and torch.equal(c["cand_cls"][:n].cpu().long(), ref_idx[:, 1])
        clipped = props.clone()
        clipped[:, 0::2] = clipped[:, 0::2].clamp(0, hw[1])
        clipped[:, 1::2] = clipped[:, 1::2].clamp(0, hw[0])
        assert torch.equal(c["cand_boxes"][:n].cpu(), clipped[ref_idx[:, 0]]), case     # small-integer boxes decode exactly
        r = O.fast_rcnn_inference_single_image(props.repeat(1, K), probs, hw, 0.05, 0.5, topk)
        k = int(out["counts"][0])
        assert k == len(r["scores"]), case
        assert torch.equal(out["roi_inds"][0, :k].cpu(), r["roi_inds"]), case
        assert torch.equal(out["classes"][0, :k].cpu(), r["classes"]), case
        assert torch.equal(out["scores"][0, :k].cpu(), r["scores"]), case
        assert torch.equal(out["boxes"][0, :k].cpu(), r["boxes"]), case


@pytest.mark.parametrize("K", [20, 80])
def test_many_cta_compaction_equals_single_cta_bitwise(K):
    """The many-CTA softmax + decode + threshold compaction (large proposal counts, BASELINE configs[4]) writes the same
    candidate list, in the same torch.nonzero() order and with the same bits, as the one-CTA-per-image kernel: ragged
    images (0, 1, 255, 256, 257, 3000, 8192 ROIs), every count checked against the probabilities' own nonzero()."""
    import ctypes
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, ops
    sizes = [0, 1, 255, 256, 257, 3000, 8192]
    gen = torch.Generator().manual_seed(K)
    R = sum(sizes)
    lg = torch.randn(R, K + 1, generator=gen)
    peak = torch.rand(R, generator=gen) < 0.3
    lg[torch.arange(R)[peak], torch.randint(0, K, (R,), generator=gen)[peak]] += 4.0
    lg[~peak, K] += 4.0
    dl = torch.randn(R, 4 * K, generator=gen) * 0.5
    pb = torch.cat([synth_proposals(max(s, 1), 600, 800, gen)[0][:s] for s in sizes], 0)
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32, device="cuda")
    hw = torch.tensor([[600.0, 800.0]] * len(sizes), device="cuda")
    lg, dl, pb = lg.cuda(), dl.cuda(), pb.cuda()
    many = ops.softmax_decode_compact(lg, dl, pb, offs, hw, 0.05, max_rois_per_image=max(sizes))
    # the single-CTA kernel through the raw C ABI (no workspace)
    cap = R * K
    one = dict(probs=torch.empty_like(lg), cand_boxes=torch.zeros(cap, 4, device="cuda"), cand_scores=torch.zeros(cap, device="cuda"),
               cand_roi=torch.zeros(cap, dtype=torch.int32, device="cuda"), cand_cls=torch.zeros(cap, dtype=torch.int32, device="cuda"),
               cand_count=torch.zeros(len(sizes), dtype=torch.int32, device="cuda"))
    _lib.call("b200_softmax_decode_compact", lg.data_ptr(), 0, dl.data_ptr(), pb.data_ptr(), offs.data_ptr(), hw.data_ptr(),
              len(sizes), R, K, 0, 10.0, 10.0, 5.0, 5.0, 0.05, one["probs"].data_ptr(), one["cand_boxes"].data_ptr(),
              one["cand_scores"].data_ptr(), one["cand_roi"].data_ptr(), one["cand_cls"].data_ptr(), one["cand_count"].data_ptr(),
              0, 0, 0, ops._stream())
    assert torch.equal(many["cand_count"], one["cand_count"]) and torch.equal(many["probs"], one["probs"])
    ref_cnt = [(int((many["probs"][int(offs[i]):int(offs[i + 1]), :K] > 0.05).sum())) for i in range(len(sizes))]
    assert many["cand_count"].tolist() == ref_cnt
    for i, s in enumerate(sizes):
        a, n = int(offs[i]) * K, ref_cnt[i]
        for key in ("cand_boxes", "cand_scores", "cand_roi", "cand_cls"):
            assert torch.equal(many[key][a:a + n], one[key][a:a + n]), (i, key)
        idx = (many["probs"][int(offs[i]):int(offs[i + 1]), :K] > 0.05).nonzero()
        assert torch.equal(many["cand_roi"][a:a + n].long(), idx[:, 0]) and torch.equal(many["cand_cls"][a:a + n].long(), idx[:, 1])
