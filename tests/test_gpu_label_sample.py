"""S1: proposal labelling + sampling kernel (csrc/label_sample.cu) vs the reference's own run and the oracle.
Bars: matched index / fg-bg label per proposal and the row counts bit-exact; sampled rows are a subset of the right
class with the reference's counts, exactly the reference's set where nothing is subsampled; the random choice is
reproducible per seed and uniform (each foreground proposal kept with probability cap / #fg, 5 sigma)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def _inputs(g, ids=(0, 1, 2, 3)):
    props = [T(g["all_props%d" % i]).cuda() for i in ids]
    gtb = [T(g["gt_boxes%d" % i]).float().reshape(-1, 4).cuda() for i in ids]
    gtc = [T(g["gt_classes%d" % i]).long().cuda() for i in ids]
    return props, gtb, gtc


def test_labels_counts_and_sets_vs_reference(golden):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    g = golden("label_sample")
    K, B = int(g["num_classes"]), int(g["batch"])
    props, gtb, gtc = _inputs(g)
    r = ops.label_and_sample_proposals(props, gtb, gtc, K, 0.5, B, float(g["pos_frac"]), seed=11, want_labels=True)
    off = np.cumsum([0] + [len(p) for p in props])
    counts = r["counts"].cpu().tolist()
    for i in range(4):
        mi, ml = r["matched_idx"][off[i]:off[i + 1]].cpu().long(), r["matched_label"][off[i]:off[i + 1]].cpu().long()
        if i != 2:
            assert torch.equal(mi, T(g["matched_idx%d" % i])) and torch.equal(ml, T(g["matched_label%d" % i]))
        else:
            assert int(ml.sum()) == 0
        ref_c = T(g["out_classes%d" % i])
        n_fg_ref, n_ref = int((ref_c < K).sum()), len(ref_c)
        assert counts[i] == [n_fg_ref, n_ref]
        n = counts[i][1]
        idx = r["sampled_idx"][i, :n].cpu().long()
        assert len(torch.unique(idx)) == n and bool((r["sampled_idx"][i, n:] == -1).all()) and bool((r["classes"][i, n:] == -1).all())
        cls, box, gbx = r["classes"][i, :n].cpu(), r["boxes"][i, :n].cpu(), r["gt_boxes"][i, :n].cpu()
        assert torch.equal(box, props[i].cpu()[idx])
        assert bool((cls[:counts[i][0]] < K).all()) and bool((cls[counts[i][0]:] == K).all())      # foreground rows first
        assert torch.equal(ml[idx], (cls < K).long())
        if i != 2:
            want_cls = torch.where(ml[idx] == 1, gtc[i].cpu()[mi[idx]], torch.full_like(cls, K))
            assert torch.equal(cls, want_cls) and torch.equal(gbx, gtb[i].cpu()[mi[idx]])
        else:
            assert float(gbx.abs().max()) == 0.0
    # image 3: fewer proposals than the batch, nothing subsampled -> exactly the reference's rows (as a set)
    n = counts[3][1]
    mine = torch.cat([r["boxes"][3, :n].cpu(), r["classes"][3, :n].cpu().float()[:, None], r["gt_boxes"][3, :n].cpu()], 1)
    ref = torch.cat([T(g["out_props3"]), T(g["out_classes3"]).float()[:, None], T(g["out_gt3"])], 1)
    key = lambda m: m[np.lexsort(m.numpy().T[::-1])]
    assert torch.equal(key(mine), key(ref))


def test_sampling_is_reproducible_and_uniform(golden):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    g = golden("label_sample")
    K, B = int(g["num_classes"]), int(g["batch"])
    props, gtb, gtc = _inputs(g, ids=(0,))
    a = ops.label_and_sample_proposals(props, gtb, gtc, K, 0.5, B, 0.25, seed=5)
    b = ops.label_and_sample_proposals(props, gtb, gtc, K, 0.5, B, 0.25, seed=5)
    c = ops.label_and_sample_proposals(props, gtb, gtc, K, 0.5, B, 0.25, seed=6)
    assert torch.equal(a["sampled_idx"], b["sampled_idx"]) and not torch.equal(a["sampled_idx"], c["sampled_idx"])
    lab = O.label_proposals(g["all_props0"], g["gt_boxes0"])[1]
    fg = torch.nonzero(lab == 1).squeeze(1)
    trials, hits = 400, torch.zeros(len(lab))
    for s in range(trials):
        r = ops.label_and_sample_proposals(props, gtb, gtc, K, 0.5, B, 0.25, seed=1000 + s)
        hits[r["sampled_idx"][0, :128].cpu().long()] += 1
    assert float(hits[lab == 0].sum()) == 0.0                               # the first 128 rows are foreground only
    p = 128.0 / len(fg)
    sigma = (trials * p * (1 - p)) ** 0.5
    assert float((hits[fg] - trials * p).abs().max()) < 5 * sigma
    assert abs(float(hits[fg].mean()) - trials * p) < 1e-3


def test_head_api_uses_the_kernel(golden):
    """ROIHeads.label_and_sample_proposals on CUDA inputs: Instances of the reference's lengths / class counts."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
    g = golden("label_sample")
    K = int(g["num_classes"])
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME, cfg.MODEL.ROI_HEADS.NUM_CLASSES = "SematicRes5ROIHeads", K
    cfg.MODEL.ADDITION.NAME = "clip"
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 4, 1
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=16, stride=16)}).cuda().train()
    assert bool(m.proposal_append_gt) == bool(int(g["append_gt"]))
    props, targets = [], []
    for i in range(4):
        p = Instances((600, 800))
        p.proposal_boxes = Boxes(T(g["props%d" % i]).cuda())
        p.objectness_logits = torch.zeros(len(g["props%d" % i]), device="cuda")
        t = Instances((600, 800))
        t.gt_boxes = Boxes(T(g["gt_boxes%d" % i]).float().reshape(-1, 4).cuda())
        t.gt_classes = T(g["gt_classes%d" % i]).long().cuda()
        props.append(p)
        targets.append(t)
    out = m.label_and_sample_proposals(props, targets)
    for i, o in enumerate(out):
        ref_c = T(g["out_classes%d" % i])
        assert len(o) == len(ref_c) and int((o.gt_classes < K).sum()) == int((ref_c < K).sum())
        assert o.gt_boxes.tensor.shape == (len(ref_c), 4) and o.proposal_boxes.tensor.shape == (len(ref_c), 4)


def test_limits_empty_image_and_ties():
    """Maximum sizes (4096 proposals, 256 ground-truth boxes per image), an image without proposals, duplicate ground-truth
    boxes (exact IoU ties resolve to the first index, as torch.max does on the CPU), and the refusal beyond the limits."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, ops
    from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals
    gen = torch.Generator().manual_seed(3)
    big, objs = synth_proposals(4096, 600, 800, gen, n_obj=128)
    gts = torch.cat([objs, objs], 0)                                   # 256 boxes, every one duplicated -> ties
    cls = torch.arange(256) % 20
    props = [big.cuda(), big[:0].cuda(), big[:700].cuda()]
    gtb = [gts.cuda(), gts[:3].cuda(), gts[:0].cuda()]
    gtc = [cls.cuda(), cls[:3].cuda(), cls[:0].cuda()]
    r = ops.label_and_sample_proposals(props, gtb, gtc, 20, 0.5, 512, 0.25, seed=1, want_labels=True)
    idx, lab, _ = O.label_proposals(big, gts)
    assert int(idx.max()) < 128                                         # first of the two identical boxes
    assert torch.equal(r["matched_idx"][:4096].cpu().long(), idx) and torch.equal(r["matched_label"][:4096].cpu().long(), lab)
    counts = r["counts"].cpu().tolist()
    n_fg = int(lab.sum())
    assert counts[0] == [min(128, n_fg), min(128, n_fg) + min(512 - min(128, n_fg), 4096 - n_fg)]
    assert counts[1] == [0, 0] and bool((r["sampled_idx"][1] == -1).all())
    assert counts[2] == [0, 512] and bool((r["classes"][2] == 20).all())
    with pytest.raises(_lib.B200Error):
        ops.label_and_sample_proposals([torch.zeros(4097, 4).cuda()], [gts[:1].cuda()], [cls[:1].cuda()], 20)
