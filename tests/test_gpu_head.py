"""Whole ROI head through the Detectron2-style API vs the reference golden (tiny config) and the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    """fp32 res5 comparisons against the CPU oracle need true fp32 convolutions (cuDNN defaults to TF32)."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old

from oracle import oracle as O
from oracle.gen_golden import synth_proposals


def T(a):
    return torch.from_numpy(np.asarray(a))


def _tiny_head(golden, tag, head, layer):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
    g = golden(tag)
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME, cfg.MODEL.ROI_HEADS.OUTPUT_LAYER = head, layer
    cfg.MODEL.ADDITION.NAME = "clip"
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 4, 1
    cfg.MODEL.B200.RES5_DTYPE = "float32"
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=16, stride=16)})
    sd = {k: T(g[k]) for k in m.state_dict().keys()}
    m.load_state_dict(sd)                                     # reference state-dict names load unchanged
    m.attention.embed, m.attention.bg_feature = T(g["embed"]), T(g["bg_feature"])
    m = m.cuda().eval()
    props = []
    for i in range(2):
        inst = Instances(tuple(int(v) for v in g["hw%d" % i]))
        inst.proposal_boxes = Boxes(T(g["props%d" % i]).cuda())
        inst.objectness_logits = torch.zeros(len(g["props%d" % i]), device="cuda")
        props.append(inst)
    return g, m, props


@pytest.mark.parametrize("tag,head,layer", [("head_tiny", "SematicRes5ROIHeads", "FastRCNNOutputLayers"),
                                            ("head_tiny_cross", "SematicRes5ROIHeadsCrossOutput", "FastRCNNAttentionOutputLayers")])
def test_head_tiny_golden(golden, tag, head, layer):
    g, m, props = _tiny_head(golden, tag, head, layer)
    feat = T(g["feat"]).cuda()
    with torch.no_grad():
        pooled = m.pooler([feat], [p.proposal_boxes for p in props])
        torch.testing.assert_close(pooled.cpu().contiguous(), T(g["pooled"]), rtol=1e-5, atol=1e-5)
        fp = m._pooled({"res4": feat}, props)
        torch.testing.assert_close(fp.cpu(), T(g["feature_pooled"]), rtol=1e-3, atol=1e-4)
        att, _ = m.forward_att(fp)
        ref_l, got_l = T(g["logits"]), att["pred_logits"].float().cpu()
        assert float((got_l - ref_l).norm() / ref_l.norm()) < 2e-2
        torch.testing.assert_close(got_l, ref_l, rtol=2e-2, atol=2e-2 * float(ref_l.abs().max()))
        ref_d, got_d = T(g["deltas"]), att["pred_bbox"].float().cpu()
        assert float((got_d - ref_d).norm() / ref_d.norm()) < 2e-2
        torch.testing.assert_close(got_d, ref_d, rtol=2e-2, atol=2e-2 * float(ref_d.abs().max()))
        res, losses = m(None, {"res4": feat}, props, None)
    assert losses == {}
    for i, r in enumerate(res):
        ref_s = T(g["det_scores%d" % i])
        assert r.image_size == tuple(int(v) for v in g["hw%d" % i])
        assert len(r) <= 100 and bool((r.scores[:-1] >= r.scores[1:]).all())
        # bf16 logits move scores by ~1e-2: compare the confident detections as sets of (class, box)
        conf = ref_s > 0.3
        ref_b, ref_c = T(g["det_boxes%d" % i])[conf], T(g["det_classes%d" % i])[conf]
        gb, gc = r.pred_boxes.tensor.cpu(), r.pred_classes.cpu()
        hit = 0
        for b, c in zip(ref_b, ref_c):
            d = (gb - b).abs().max(dim=1).values
            hit += bool(((d < 2.0) & (gc == c)).any())
        assert hit >= 0.9 * len(ref_b)


def test_cross_output_cosine_logits_option(golden):
    """MODEL.B200.COSINE_LOGITS: the CrossOutput head's prototype logits become tau * cos(projected feature, text
    prototype) (the reference's `sim_matrix` semantics, my_module.py:461-469); kernel path (eval) vs the torch expression
    of the training branch on the same weights, bf16 bar.  Default off = the golden (un-normalised) logits."""
    g, m, props = _tiny_head(golden, "head_tiny_cross", "SematicRes5ROIHeadsCrossOutput", "FastRCNNAttentionOutputLayers")
    assert m.cosine_logits is False
    feat = T(g["feat"]).cuda()
    m.cosine_logits, m.cosine_tau = True, 10.0
    with torch.no_grad():
        fp = m._pooled({"res4": feat}, props)
        got = m.forward_att(fp)[0]["pred_logits"].float()
        _, oa = m.attention(fp)
        a = F.relu(m.output_projection(oa["sim2stext"].float()))
        want = O.sim_matrix(a.cpu(), oa["text_feat"].float().cpu(), tau=10.0)
    assert got.shape == want.shape and float(want.abs().max()) <= 10.0 + 1e-3
    assert float((got.cpu() - want).norm() / want.norm()) < 2e-2


def test_head_full_width_vs_oracle():
    """Real channel widths (1024 -> 2048), R = 2 x 48, fp32 res5: logits vs the fp32 CPU oracle within the bf16 bar."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = "SematicRes5ROIHeads"
    cfg.MODEL.ADDITION.NAME = "clip"
    cfg.MODEL.B200.RES5_DTYPE = "float32"
    torch.manual_seed(0)
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)}).eval()
    with torch.no_grad():
        m.box_predictor.cls_score.weight.mul_(30.0)
        m.box_predictor.bbox_pred.weight.mul_(50.0)
    gen = torch.Generator().manual_seed(4)
    feat = torch.relu(torch.randn(2, 1024, 25, 32, generator=gen)) * 0.5
    sizes = [(400, 512), (384, 500)]
    boxes = [synth_proposals(48, h, w, gen)[0] for (h, w) in sizes]
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    text = torch.cat([m.attention.embed, m.attention.bg_feature], 0)
    dets_ref, mid = O.head_forward(feat, boxes, sizes, text, p)
    m = m.cuda()
    props = []
    for b, s in zip(boxes, sizes):
        inst = Instances(s)
        inst.proposal_boxes = Boxes(b.cuda())
        props.append(inst)
    with torch.no_grad():
        fp = m._pooled({"res4": feat.cuda()}, props)
        att, _ = m.forward_att(fp)
        res, _ = m(None, {"res4": feat.cuda()}, props, None)
    torch.testing.assert_close(fp.cpu(), mid["feature_pooled"], rtol=2e-3, atol=2e-3)
    rl = mid["logits"]
    assert float((att["pred_logits"].cpu() - rl).norm() / rl.norm()) < 2e-2
    torch.testing.assert_close(att["pred_logits"].cpu().float(), rl, rtol=2e-2, atol=2e-2 * float(rl.abs().max()))
    rd = mid["deltas"]
    assert float((att["pred_bbox"].cpu() - rd).norm() / rd.norm()) < 2e-2
    torch.testing.assert_close(att["pred_bbox"].cpu().float(), rd, rtol=2e-2, atol=2e-2 * float(rd.abs().max()))
    torch.testing.assert_close(att["sim2stext"].cpu().float(), mid["sim2stext"], rtol=2e-2, atol=2e-2 * float(mid["sim2stext"].abs().max()))
    assert len(res) == 2 and all(len(r) <= 100 for r in res)


def test_head_full_width_bf16_res5_vs_oracle():
    """The BENCH path at real channel widths: bf16 res5 on the own tcgen05 kernels (the default), R = 2 x 48, against the
    fp32 CPU oracle — pooled feature, fused feature, logits and deltas elementwise within north_star's bf16 bar
    (|got - ref| <= 2e-2 |ref| + 2e-2 max |ref|) and in relative L2."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = "SematicRes5ROIHeads"
    cfg.MODEL.ADDITION.NAME = "clip"
    assert cfg.MODEL.B200.RES5_DTYPE == "bfloat16" and cfg.MODEL.B200.RES5_IMPL == "tcgen05"
    torch.manual_seed(0)
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)}).eval()
    with torch.no_grad():
        m.box_predictor.cls_score.weight.mul_(30.0)
        m.box_predictor.bbox_pred.weight.mul_(50.0)
    gen = torch.Generator().manual_seed(4)
    feat = torch.relu(torch.randn(2, 1024, 25, 32, generator=gen)) * 0.5
    sizes = [(400, 512), (384, 500)]
    boxes = [synth_proposals(48, h, w, gen)[0] for (h, w) in sizes]
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    text = torch.cat([m.attention.embed, m.attention.bg_feature], 0)
    _, mid = O.head_forward(feat, boxes, sizes, text, p)
    m = m.cuda()
    props = []
    for b, s in zip(boxes, sizes):
        inst = Instances(s)
        inst.proposal_boxes = Boxes(b.cuda())
        props.append(inst)
    with torch.no_grad():
        fp = m._pooled({"res4": feat.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)}, props)
        att, _ = m.forward_att(fp)
        res, _ = m(None, {"res4": feat.cuda()}, props, None)

    def close(got, ref):
        got, ref = got.float().cpu(), ref.float()
        assert float((got - ref).norm() / ref.norm()) < 2e-2
        torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    close(fp, mid["feature_pooled"])
    close(att["sim2stext"], mid["sim2stext"])
    close(att["pred_logits"], mid["logits"])
    close(att["pred_bbox"], mid["deltas"])
    assert len(res) == 2 and all(len(r) <= 100 for r in res)


@pytest.mark.parametrize("head_name", ["SematicRes5ROIHeads", "SematicRes5ROIHeadsDistill"])
def test_training_step_runs_and_backprops(head_name):
    """Fine-tune step through the Detectron2-style forward: losses (loss_cls, loss_box_reg, loss_attentive; + loss_kl for
    the distillation head, BASELINE configs[3]) and gradients reach the res4 map through the ROIAlign backward kernel and
    the fused GDL/affine backward."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = head_name
    cfg.MODEL.ADDITION.NAME = "clip"
    cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE = 64
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 16, 4
    torch.manual_seed(0)
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=64, stride=16)}).cuda().train()
    for p_ in m.res5.parameters():
        p_.requires_grad = False
    aff = modeling.AffineLayer(64, bias=True).cuda()
    gen = torch.Generator().manual_seed(1)
    base = torch.relu(torch.randn(2, 64, 20, 25, generator=gen)).cuda().requires_grad_(True)
    feat = modeling.decoupled_affine(base, aff, 0.01, channels_last_out=True)
    props, tgts = [], []
    for _ in range(2):
        b, objs = synth_proposals(100, 320, 400, gen)
        inst = Instances((320, 400))
        inst.proposal_boxes = Boxes(b.cuda())
        inst.objectness_logits = torch.zeros(100, device="cuda")
        props.append(inst)
        t = Instances((320, 400))
        t.gt_boxes = Boxes(objs.cuda())
        t.gt_classes = torch.randint(0, 20, (len(objs),), generator=gen).cuda()
        tgts.append(t)
    _, losses = m(None, {"res4": feat}, props, tgts)
    want = {"loss_cls", "loss_box_reg", "loss_attentive"} | ({"loss_kl"} if head_name.endswith("Distill") else set())
    assert set(losses) == want and all(torch.isfinite(v) for v in losses.values())
    if head_name.endswith("Distill"):
        assert float(losses["loss_kl"]) >= 0 and all(p_.grad is None for p_ in m.teacher.parameters())
    sum(losses.values()).backward()
    assert base.grad is not None and torch.isfinite(base.grad).all() and float(base.grad.abs().sum()) > 0
    assert aff.weight.grad is not None and torch.isfinite(aff.weight.grad).all()
    assert m.attention.attention.w_q.weight.grad is not None


def test_end_to_end_voc_ap_within_0p1():
    """north_star: end-to-end VOC-style AP within 0.1.  Proposals -> head (bf16 fusion chain, CUDA post-processing) ->
    detector_postprocess -> VOC text lines -> AP, against the fp32 CPU oracle taken through the same records.  Ground
    truth = the oracle's own confident detections (threshold placed in the widest score gap), so the oracle scores
    AP = 100 and every ranking flip, lost or spurious detection of the CUDA path costs AP."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling, ops
    from fewshotobjectdetection_imporove_via_text_feature_b200.evaluation import detection_formats as DF
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.postprocessing import detector_postprocess_batch
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = "SematicRes5ROIHeads"
    cfg.MODEL.ADDITION.NAME = "clip"
    cfg.MODEL.B200.RES5_DTYPE = "float32"
    torch.manual_seed(0)
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)}).eval()
    gen = torch.Generator().manual_seed(12)
    sizes = [(400, 512), (384, 500), (416, 480)]
    out_sizes = [(333, 426), (384, 500), (832, 960)]
    feat = torch.relu(torch.randn(3, 1024, 26, 32, generator=gen)) * 0.5
    boxes = [synth_proposals(64, h, w, gen)[0] for (h, w) in sizes]
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    text = torch.cat([m.attention.embed, m.attention.bg_feature], 0)
    _, mid = O.head_forward(feat, boxes, sizes, text, p)
    # a random-init classifier puts every ROI in one class: sharpen it and centre its logits on the mean fused feature,
    # so that the detections spread over (nearly) all 20 classes with scores across (0, 1)
    with torch.no_grad():
        m.box_predictor.cls_score.weight.mul_(60.0)
        m.box_predictor.cls_score.bias.copy_(-(m.box_predictor.cls_score.weight @ mid["sim2stext"].mean(0)))
        m.box_predictor.bbox_pred.weight.mul_(50.0)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    logits, deltas = O.output_layers(mid["feature_pooled"], mid["sim2stext"], p)
    dets_ref = O.fast_rcnn_inference(logits, deltas, boxes, sizes, 0.05, 0.5, 100)
    ids = ["img%d" % i for i in range(3)]

    def records(b, s, c, n):
        lines = DF.voc_prediction_lines(ids, b, s, c, n)
        return {cls: [(l.split()[0],) + tuple(float(v) for v in l.split()[1:]) for l in ls] for cls, ls in lines.items()}

    # oracle side: postprocess + the same record format
    T_ = 100
    rb, rs_, rc, rn = np.zeros((3, T_, 4), np.float32), np.zeros((3, T_), np.float32), np.zeros((3, T_), np.int64), np.zeros(3, np.int64)
    for i, d in enumerate(dets_ref):
        bb, keep = O.detector_postprocess(d["boxes"], sizes[i], *out_sizes[i])
        k = len(bb)
        rb[i, :k], rs_[i, :k], rc[i, :k], rn[i] = bb.numpy(), d["scores"][keep].numpy(), d["classes"][keep].numpy(), k
    ref = records(rb, rs_, rc, rn)
    all_scores = np.sort(np.concatenate([rs_[i, :rn[i]] for i in range(3)]))
    band = all_scores[(all_scores > 0.2) & (all_scores < 0.8)]
    assert len(band) >= 2
    gi = int(np.argmax(np.diff(band)))
    tau = 0.5 * (band[gi] + band[gi + 1])
    gts = {}
    for cls, ds in ref.items():
        for (img, sc, x1, y1, x2, y2) in ds:
            if sc > tau:
                gts.setdefault(cls, {}).setdefault(img, []).append([x1, y1, x2, y2])
    assert len(gts) >= 8, "degenerate synthetic case: too few confident classes"

    # CUDA side
    m = m.cuda()
    props = []
    for b, s in zip(boxes, sizes):
        inst = Instances(s)
        inst.proposal_boxes = Boxes(b.cuda())
        props.append(inst)
    with torch.no_grad():
        fp = m._pooled({"res4": feat.cuda()}, props)
        att, _ = m.forward_att(fp)
        offs = torch.tensor([0, 64, 128, 192], dtype=torch.int32, device="cuda")
        det = ops.fast_rcnn_inference_device(att["pred_logits"], att["pred_bbox"], torch.cat(boxes).cuda(), offs,
                                             ops.image_hw_tensor(sizes, "cuda"), 0.05, 0.5, 100)
    detector_postprocess_batch(det, sizes, out_sizes)
    got = records(*DF.pack_batch(det))

    aps_ref, aps_got = [], []
    for cls, g in gts.items():
        g = {k: np.array(v) for k, v in g.items()}
        aps_ref.append(O.voc_eval_class(ref.get(cls, []), g) * 100)
        aps_got.append(O.voc_eval_class(got.get(cls, []), g) * 100)
    assert min(aps_ref) > 99.999                         # the oracle against its own confident detections
    assert abs(np.mean(aps_got) - np.mean(aps_ref)) <= 0.1, (aps_got, aps_ref)
