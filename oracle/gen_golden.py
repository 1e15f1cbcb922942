"""oracle/gen_golden.py — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the reference's OWN modules, imported unchanged from
/root/reference behind oracle/ref_stubs.py, on small seeded inputs.  Run in the authoring container
(where /root/reference is mounted); the resulting vectors are committed so that the GPU box — where
the reference does not exist — can still check the oracle port (and through it the CUDA path).

    python -m oracle.gen_golden            # writes tests/golden/
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

from . import ref_stubs as rs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@contextlib.contextmanager
def cuda_as_cpu():
    """The reference hard-codes device='cuda' (attentive_modules.py:199, class_embedding.py:13,
    roi_heads.py:927).  Redirect those `.to('cuda')` calls to CPU without touching the sources."""
    orig = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return orig(self, *a, **k)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


def sd_to_np(sd, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals, synth_rpn_outputs  # noqa: E402,F401


def gen_gdl():
    gdl = rs.load("defrcn.modeling.meta_arch.gdl")
    torch.manual_seed(0)
    x = torch.relu(torch.randn(2, 6, 5, 7)).requires_grad_(True)
    aff = gdl.AffineLayer(6, bias=True)
    with torch.no_grad():
        aff.weight.copy_(torch.randn(1, 6, 1, 1))
        aff.bias.copy_(torch.randn(1, 6, 1, 1))
    lam = 0.001
    y = aff(gdl.decouple_layer(x, lam))
    g = torch.randn_like(y)
    y.backward(g)
    np.savez_compressed(os.path.join(OUT, "gdl.npz"), x=x.detach().numpy(), w=aff.weight.detach().numpy(),
                        b=aff.bias.detach().numpy(), lam=np.float32(lam), y=y.detach().numpy(),
                        g=g.numpy(), gx=x.grad.numpy(), gw=aff.weight.grad.numpy(), gb=aff.bias.grad.numpy())


def gen_fast_rcnn_inference():
    fr = rs.load("defrcn.modeling.roi_heads.fast_rcnn")
    gen = torch.Generator().manual_seed(11)
    out = {}
    for tag, (R, K, h, w, peaked, ties) in {
        "voc": (300, 20, 600, 800, 0.3, False),
        "coco": (200, 80, 480, 640, 0.5, False),
        "ties": (128, 5, 300, 400, 0.6, True),
        "empty": (16, 20, 600, 800, 0.0, False),
    }.items():
        props, _ = synth_proposals(R, h, w, gen)
        logits = torch.randn(R, K + 1, generator=gen)
        logits[:, K] += 4.0 if peaked > 0 else 12.0
        npk = int(peaked * R)
        if npk:
            cls = torch.randint(0, K, (npk,), generator=gen)
            logits[torch.arange(npk), cls] += 8.0
        if ties:  # exactly duplicated rows: equal scores and identical boxes
            logits[R // 2:] = logits[:R - R // 2]
            props[R // 2:] = props[:R - R // 2]
        deltas = torch.randn(R, 4 * K, generator=gen) * 0.5
        if ties:
            deltas[R // 2:] = deltas[:R - R // 2]
        inst = rs.Instances((h, w))
        inst.proposal_boxes = rs.Boxes(props.clone())
        o = fr.FastRCNNOutputs(rs.Box2BoxTransform((10.0, 10.0, 5.0, 5.0)), logits, deltas, [inst], 0.0)
        boxes = o.predict_boxes()[0]
        probs = o.predict_probs()[0]
        res, roi_inds = fr.fast_rcnn_inference_single_image(boxes.clone(), probs, (h, w), 0.05, 0.5, 100)
        ncand = int((probs[:, :-1] > 0.05).sum())
        out.update({
            tag + "_props": props.numpy(), tag + "_logits": logits.numpy(), tag + "_deltas": deltas.numpy(),
            tag + "_hw": np.array([h, w], np.int64), tag + "_pred_boxes_all": boxes.numpy(),
            tag + "_probs": probs.numpy(), tag + "_ncand": np.int64(ncand),
            tag + "_boxes": res.pred_boxes.tensor.numpy(), tag + "_scores": res.scores.numpy(),
            tag + "_classes": res.pred_classes.numpy(), tag + "_roi_inds": roi_inds.numpy(),
        })
    np.savez_compressed(os.path.join(OUT, "fast_rcnn_inference.npz"), **out)


def gen_losses():
    fr = rs.load("defrcn.modeling.roi_heads.fast_rcnn")
    gen = torch.Generator().manual_seed(5)
    R, K, h, w = 96, 15, 600, 800
    props, objs = synth_proposals(R, h, w, gen)
    gt_classes = torch.randint(0, K + 1, (R,), generator=gen)
    gt_classes[R // 3:] = K
    gt_boxes = props + torch.randn(R, 4, generator=gen) * 4
    gt_boxes[:, 2:] = torch.maximum(gt_boxes[:, 2:], gt_boxes[:, :2] + 2)
    logits = torch.randn(R, K + 1, generator=gen)
    deltas = torch.randn(R, 4 * K, generator=gen) * 0.1
    inst = rs.Instances((h, w))
    inst.proposal_boxes = rs.Boxes(props)
    inst.gt_boxes = rs.Boxes(gt_boxes)
    inst.gt_classes = gt_classes
    o = fr.FastRCNNOutputs(rs.Box2BoxTransform((10.0, 10.0, 5.0, 5.0)), logits, deltas, [inst], 0.0)
    L = o.losses()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), props=props.numpy(), gt_boxes=gt_boxes.numpy(),
                        gt_classes=gt_classes.numpy(), logits=logits.numpy(), deltas=deltas.numpy(),
                        loss_cls=L["loss_cls"].numpy(), loss_box_reg=L["loss_box_reg"].numpy(),
                        cls_accuracy=np.float64(rs.get_event_storage().scalars["fast_rcnn/cls_accuracy"]))


def gen_attention(d_model=32, R=40, K=20, tag="attention_small"):
    rs.install(class_embed_fn=lambda names, model, include_bg=False: rs.synthetic_class_embed(names, model, include_bg))
    am = rs.load("defrcn.modeling.roi_heads.attentive_modules")
    am.get_class_embed = lambda names, model, include_bg=False: rs.synthetic_class_embed(names, model, include_bg)
    cfg = rs.default_cfg(num_classes=K, addition="clip")
    torch.manual_seed(0)
    with cuda_as_cpu():
        att = am.SematicProposalAttention(d_model, cfg=cfg).eval()
        x = torch.relu(torch.randn(R, d_model, generator=torch.Generator().manual_seed(3)))
        with torch.no_grad():
            attn, out = att(x)
    d = sd_to_np(att.state_dict(), "attention.")
    d.update(x=x.numpy(), embed=att.embed.numpy(), bg_feature=att.bg_feature.numpy(),
             sim2stext=out["sim2stext"].numpy(), text_feat=out["text_feat"].numpy(), attn=attn[0].numpy())
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **d)


def gen_head_tiny():
    """Whole SematicRes5ROIHeads / SematicRes5ROIHeadsCrossOutput eval forward on a shrunken config
    (res4 channels 16 -> res5 32) so the weights fit in a fixture."""
    emb = lambda names, model, include_bg=False: rs.synthetic_class_embed(names, model, include_bg)
    rs.install(class_embed_fn=emb)
    am = rs.load("defrcn.modeling.roi_heads.attentive_modules")
    am.get_class_embed = emb
    rh = rs.load("defrcn.modeling.roi_heads.roi_heads")
    for head, layer, tag in (("SematicRes5ROIHeads", "FastRCNNOutputLayers", "head_tiny"),
                             ("SematicRes5ROIHeadsCrossOutput", "FastRCNNAttentionOutputLayers", "head_tiny_cross")):
        K = 20
        cfg = rs.default_cfg(num_classes=K, addition="clip", output_layer=layer, roi_head=head)
        cfg.MODEL.RESNETS.RES2_OUT_CHANNELS = 4      # res5 out = 32, in = 16
        cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 1        # bottleneck = 8
        torch.manual_seed(1)
        with cuda_as_cpu():
            m = rh.build_roi_heads(cfg, {"res4": rs.ShapeSpec(channels=16, stride=16)}).eval()
            with torch.no_grad():
                # make the classifier outputs non-trivial (reference init std=0.01/0.001 gives ~uniform probs)
                m.box_predictor.cls_score.weight.mul_(60.0)
                m.box_predictor.bbox_pred.weight.mul_(100.0)
                for blk in m.res5:
                    for c in (blk.conv1, blk.conv2, blk.conv3, blk.shortcut):
                        if c is not None:
                            c.norm.weight.copy_(torch.rand_like(c.norm.weight) + 0.5)
                            c.norm.bias.copy_(torch.randn_like(c.norm.bias) * 0.1)
                            c.norm.running_mean.copy_(torch.randn_like(c.norm.running_mean) * 0.1)
                            c.norm.running_var.copy_(torch.rand_like(c.norm.running_var) + 0.5)
            gen = torch.Generator().manual_seed(21)
            sizes = [(320, 400), (304, 464)]
            Hf, Wf = 20, 29
            feat = torch.relu(torch.randn(2, 16, Hf, Wf, generator=gen))
            props = []
            for (h, w) in sizes:
                b, _ = synth_proposals(48, h, w, gen)
                inst = rs.Instances((h, w))
                inst.proposal_boxes = rs.Boxes(b)
                inst.objectness_logits = torch.zeros(len(b))
                props.append(inst)
            with torch.no_grad():
                res, _ = m(None, {"res4": feat}, props, None)
                pooled = m.pooler([feat], [p.proposal_boxes for p in props])
                fp = m.res5(pooled).mean(dim=[2, 3])
                att_out, _ = m.forward_att(fp)
        d = sd_to_np(m.state_dict())
        d.update(feat=feat.numpy(), embed=m.attention.embed.numpy(), bg_feature=m.attention.bg_feature.numpy(),
                 pooled=pooled.numpy(), feature_pooled=fp.numpy(), logits=att_out["pred_logits"].numpy(),
                 deltas=att_out["pred_bbox"].numpy(), sim2stext=att_out["sim2stext"].numpy())
        for i, (p, r) in enumerate(zip(props, res)):
            d["props%d" % i] = p.proposal_boxes.tensor.numpy()
            d["hw%d" % i] = np.array(p.image_size, np.int64)
            d["det_boxes%d" % i] = r.pred_boxes.tensor.numpy()
            d["det_scores%d" % i] = r.scores.numpy()
            d["det_classes%d" % i] = r.pred_classes.numpy()
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **d)


def gen_pcb():
    """Runs PrototypicalCalibrationBlock.execute_calibration / extract_roi_features unchanged, with the
    ImageNet CNN replaced by a fixed random conv feature (the CNN itself is out of scope)."""
    rs.install()
    sys.modules.setdefault("defrcn.dataloader", types.ModuleType("defrcn.dataloader"))
    sys.modules["defrcn.dataloader"].build_detection_test_loader = None
    archs = types.ModuleType("defrcn.evaluation.archs")
    archs.resnet101 = None
    sys.modules["defrcn.evaluation.archs"] = archs

    def from_tensors(tensors, size_divisibility=0):
        return rs.ImageList(torch.stack(tensors), [t.shape[-2:] for t in tensors])
    rs.ImageList.from_tensors = staticmethod(from_tensors)
    cl = rs.load("defrcn.evaluation.calibration_layer")
    gen = torch.Generator().manual_seed(9)
    h, w, D, K, n = 416, 608, 48, 20, 60
    conv_feature = torch.relu(torch.randn(1, 64, h // 32, w // 32, generator=gen))

    class FakeNet:
        fc = torch.nn.Linear(64, D)

        def __call__(self, x):
            return None, conv_feature
    torch.manual_seed(4)
    net = FakeNet()
    pcb = object.__new__(cl.PrototypicalCalibrationBlock)
    pcb.cfg = rs.default_cfg()
    pcb.device = torch.device("cpu")
    pcb.alpha = 0.5
    pcb.imagenet_model = net
    pcb.roi_pooler = rs.ROIPooler(output_size=(1, 1), scales=(1 / 32,), sampling_ratio=(0), pooler_type="ROIAlignV2")
    protos = torch.randn(K, D, generator=gen)
    pcb.prototypes = {c: protos[c:c + 1] for c in range(K)}
    pcb.exclude_cls = list(range(0, 15))
    boxes, _ = synth_proposals(n, h, w, gen)
    scores = torch.sort(torch.rand(n, generator=gen) * 1.02, descending=True).values
    scores[-7:] = scores[-7:] * 0.04           # a tail below PCB_LOWER
    classes = torch.randint(0, K, (n,), generator=gen)
    inst = rs.Instances((h, w))
    inst.pred_boxes = rs.Boxes(boxes.clone())
    inst.scores = scores.clone()
    inst.pred_classes = classes
    img = np.zeros((h, w, 3), np.uint8)
    cl.cv2.imread = lambda fn: img
    with torch.no_grad():
        feats = pcb.extract_roi_features(img, [inst.pred_boxes])
        dts = pcb.execute_calibration([{"file_name": "x"}], [{"instances": inst}])
    np.savez_compressed(os.path.join(OUT, "pcb.npz"), conv_feature=conv_feature.numpy(), boxes=boxes.numpy(),
                        scores_in=scores.numpy(), classes=classes.numpy(), protos=protos.numpy(),
                        fc_w=net.fc.weight.detach().numpy(), fc_b=net.fc.bias.detach().numpy(),
                        feats_all=feats.numpy(), scores_out=dts[0]["instances"].scores.numpy(),
                        exclude=np.array(pcb.exclude_cls, np.int64), alpha=np.float32(0.5))


def seeded_fill(module, seed, scale=0.05):
    """Overwrite every parameter/buffer with values that depend only on (seed, sorted key index, shape): lets a
    golden produced by the reference module be replayed on the mirror module without storing the weights — and
    fails loudly if the two state dicts differ in names or shapes."""
    sd = module.state_dict()
    with torch.no_grad():
        for i, k in enumerate(sorted(sd)):
            g = torch.Generator().manual_seed(seed * 1000 + i)
            v = torch.randn(sd[k].shape, generator=g) * scale
            if k.endswith("norm3.weight"):
                v = v + 1.0
            sd[k].copy_(v)
    return sorted((k, tuple(v.shape)) for k, v in sd.items())


def gen_teacher():
    """KD losses (my_module.py:393-437) and the LV / textDomination teacher attentions (attentive_modules.py:297-437,
    490-634) run unchanged.  (The *_VKV forwards are broken as checked in — SURVEY §2.2 — so they have no golden.)"""
    rs.install()
    mm = rs.load("defrcn.modeling.roi_heads.my_module")
    gen = torch.Generator().manual_seed(13)
    R, K = 64, 15
    s_out, t_out = torch.randn(R, K + 1, generator=gen), torch.randn(R, K + 1, generator=gen) * 2
    labels = torch.randint(0, K + 1, (R,), generator=gen)
    labels[R // 2:] = K
    params = {"alpha": 0.7, "temperature": 5.0}
    d = dict(s_out=s_out.numpy(), t_out=t_out.numpy(), labels=labels.numpy(), alpha=np.float32(0.7), T=np.float32(5.0),
             kd=mm.loss_fn_kd(s_out, labels, t_out, params).numpy(),
             kd_only=mm.loss_fn_kd_only(s_out, labels, K, t_out, params).numpy())
    am = rs.load("defrcn.modeling.roi_heads.attentive_modules")
    cfg = rs.default_cfg(num_classes=20, addition="glove")
    cfg.MODEL.ROI_HEADS.DISTILLATE = False
    cfg.MODEL.ROI_HEADS.STUDENT_TRAINING = False
    cfg.MODEL.ROI_HEADS.TEACHER_TRAINING = True
    dm, Rr = 32, 24
    x = torch.relu(torch.randn(Rr, dm, generator=gen))
    lab = torch.randint(0, 21, (Rr,), generator=gen)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with cuda_as_cpu():
            for cls, tag in ((am.LV_attention, "lv"), (am.LV_attention_textDomination, "td")):
                torch.manual_seed(2)
                m = cls(dm, cfg=cfg).eval()
                keys = seeded_fill(m, 77)
                with torch.no_grad():
                    _, out = m(x, lab)
                d[tag + "_keys"] = np.array(["%s:%s" % (k, "x".join(map(str, shp))) for k, shp in keys])
                d[tag + "_embed"] = m.embed.numpy()
                d[tag + "_sim2stext"] = out["sim2stext"].numpy()
    finally:
        torch.Tensor.cuda = orig_cuda
    d.update(x=x.numpy(), lab=lab.numpy())
    np.savez_compressed(os.path.join(OUT, "teacher.npz"), **d)


def gen_train_step():
    """Fine-tune direction pinned on the reference's own autograd: SematicRes5ROIHeads.forward_att +
    FastRCNNOutputs.losses in train mode (roi_heads.py:1072-1132, dropout off so the numbers are reproducible) on
    fixed pooled features / labels; stores the three losses and d(sum of losses)/d(feature, every trained
    parameter of the attention and the predictor)."""
    emb = lambda names, model, include_bg=False: rs.synthetic_class_embed(names, model, include_bg)
    rs.install(class_embed_fn=emb)
    am = rs.load("defrcn.modeling.roi_heads.attentive_modules")
    am.get_class_embed = emb
    rh = rs.load("defrcn.modeling.roi_heads.roi_heads")
    fr = rs.load("defrcn.modeling.roi_heads.fast_rcnn")
    K, R = 20, 96
    cfg = rs.default_cfg(num_classes=K, addition="clip", output_layer="FastRCNNOutputLayers", roi_head="SematicRes5ROIHeads")
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS = 8      # res5 out = 64
    cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 1
    cfg.MODEL.ROI_HEADS.CLS_DROPOUT = False
    torch.manual_seed(11)
    with cuda_as_cpu():
        m = rh.build_roi_heads(cfg, {"res4": rs.ShapeSpec(channels=32, stride=16)}).train()
        with torch.no_grad():
            m.box_predictor.cls_score.weight.mul_(30.0)
            m.box_predictor.bbox_pred.weight.mul_(100.0)
            for p_ in m.attention.parameters():     # the 0.02-std init gives near-zero gradients: widen it
                p_.mul_(4.0)
        gen = torch.Generator().manual_seed(12)
        d = m.out_channels
        x = torch.relu(torch.randn(R, d, generator=gen)).requires_grad_(True)
        h, w = 600, 800
        props, _ = synth_proposals(R, h, w, gen)
        gt_classes = torch.randint(0, K + 1, (R,), generator=gen)
        gt_classes[R // 3:] = K
        gt_boxes = props + torch.randn(R, 4, generator=gen) * 4
        gt_boxes[:, 2:] = torch.maximum(gt_boxes[:, 2:], gt_boxes[:, :2] + 2)
        inst = rs.Instances((h, w))
        inst.proposal_boxes = rs.Boxes(props)
        inst.gt_boxes = rs.Boxes(gt_boxes)
        inst.gt_classes = gt_classes
        att_output, att_loss = m.forward_att(x, gt_classes)
        o = fr.FastRCNNOutputs(m.box2box_transform, att_output["pred_logits"], att_output["pred_bbox"], [inst], m.smooth_l1_beta)
        L = dict(o.losses())
        L.update(att_loss)
        sum(L.values()).backward()
    out = {k: v for k, v in sd_to_np(m.state_dict()).items() if not k.startswith("res5.")}
    out.update(x=x.detach().numpy(), props=props.numpy(), gt_boxes=gt_boxes.numpy(), gt_classes=gt_classes.numpy(),
               embed=m.attention.embed.numpy(), bg_feature=m.attention.bg_feature.numpy(), grad_x=x.grad.numpy(),
               **{"loss." + k: v.detach().numpy() for k, v in L.items()})
    for k, p_ in m.named_parameters():
        if p_.grad is not None:
            out["grad." + k] = p_.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **out)


def gen_train_step_cross():
    """C1' in the fine-tune direction, pinned on the reference's own modules: SematicRes5ROIHeadsCrossOutput.forward_att
    (roi_heads.py:1154-1171: relu(output_projection(sim2stext)) . T^T as the class logits) + FastRCNNOutputs.losses
    (fast_rcnn.py:222-304) in train mode and their autograd.  The head's own `forward` cannot run in training as checked in
    (it calls `.items()` on the attention tensor forward_att returns in place of a loss dict), so the fixture stops where the
    reference still runs: the two losses of the predictor."""
    emb = lambda names, model, include_bg=False: rs.synthetic_class_embed(names, model, include_bg)
    rs.install(class_embed_fn=emb)
    am = rs.load("defrcn.modeling.roi_heads.attentive_modules")
    am.get_class_embed = emb
    rh = rs.load("defrcn.modeling.roi_heads.roi_heads")
    fr = rs.load("defrcn.modeling.roi_heads.fast_rcnn")
    K, R = 20, 96
    cfg = rs.default_cfg(num_classes=K, addition="clip", output_layer="FastRCNNAttentionOutputLayers",
                         roi_head="SematicRes5ROIHeadsCrossOutput")
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS = 8      # res5 out = 64
    cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 1
    cfg.MODEL.ROI_HEADS.CLS_DROPOUT = False
    torch.manual_seed(21)
    with cuda_as_cpu():
        m = rh.build_roi_heads(cfg, {"res4": rs.ShapeSpec(channels=32, stride=16)}).train()
        with torch.no_grad():
            m.box_predictor.bbox_pred.weight.mul_(100.0)
            m.output_projection.weight.mul_(6.0)
            for p_ in m.attention.parameters():     # the 0.02-std init gives near-zero gradients: widen it
                p_.mul_(4.0)
        gen = torch.Generator().manual_seed(22)
        d = m.out_channels
        x = torch.relu(torch.randn(R, d, generator=gen)).requires_grad_(True)
        h, w = 600, 800
        props, _ = synth_proposals(R, h, w, gen)
        gt_classes = torch.randint(0, K + 1, (R,), generator=gen)
        gt_classes[R // 3:] = K
        gt_boxes = props + torch.randn(R, 4, generator=gen) * 4
        gt_boxes[:, 2:] = torch.maximum(gt_boxes[:, 2:], gt_boxes[:, :2] + 2)
        inst = rs.Instances((h, w))
        inst.proposal_boxes = rs.Boxes(props)
        inst.gt_boxes = rs.Boxes(gt_boxes)
        inst.gt_classes = gt_classes
        att_output, _attn = m.forward_att(x, gt_classes)
        o = fr.FastRCNNOutputs(m.box2box_transform, att_output["pred_logits"], att_output["pred_bbox"], [inst], m.smooth_l1_beta)
        L = dict(o.losses())
        sum(L.values()).backward()
    out = {k: v for k, v in sd_to_np(m.state_dict()).items() if not k.startswith("res5.")}
    out.update(x=x.detach().numpy(), props=props.numpy(), gt_boxes=gt_boxes.numpy(), gt_classes=gt_classes.numpy(),
               embed=m.attention.embed.numpy(), bg_feature=m.attention.bg_feature.numpy(), grad_x=x.grad.numpy(),
               text_feat=att_output["text_feat"].detach().numpy(), pred_logits=att_output["pred_logits"].detach().numpy(),
               **{"loss." + k: v.detach().numpy() for k, v in L.items()})
    for k, p_ in m.named_parameters():
        if p_.grad is not None:
            out["grad." + k] = p_.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "train_step_cross.npz"), **out)


def gen_known_answer():
    """test.py:80-92 fixture: CE(pred_logits.pt, gt_classes.pt)."""
    pl = torch.load(os.path.join(rs.REFERENCE_ROOT, "pred_logits.pt"), map_location="cpu").detach()
    gt = torch.load(os.path.join(rs.REFERENCE_ROOT, "gt_classes.pt"), map_location="cpu")
    ce = torch.nn.functional.cross_entropy(pl.float(), gt.long())
    np.savez_compressed(os.path.join(OUT, "known_answer_ce.npz"), ce=np.float64(ce.item()),
                        shape=np.array(pl.shape), n_bg=np.int64((gt == pl.shape[1] - 1).sum()),
                        logits_head=pl[:64].float().numpy(), gt_head=gt[:64].numpy())


def gen_label_sample():
    """The reference's own ROIHeads.label_and_sample_proposals (roi_heads.py:157-250) on three synthetic images: many
    proposals around a few objects, an image with more foreground than the 128 cap, and an image without ground truth.
    Matching / labels are deterministic; of the random subsample the fixture keeps what is: the counts, and the sampled
    foreground set where it is not subsampled."""
    emb = lambda names, model, include_bg=False: rs.synthetic_class_embed(names, model, include_bg)
    rs.install(class_embed_fn=emb)
    am = rs.load("defrcn.modeling.roi_heads.attentive_modules")
    am.get_class_embed = emb
    rh = rs.load("defrcn.modeling.roi_heads.roi_heads")
    K = 20
    cfg = rs.default_cfg(num_classes=K, addition="clip", output_layer="FastRCNNOutputLayers", roi_head="SematicRes5ROIHeads")
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 4, 1
    torch.manual_seed(1)
    with cuda_as_cpu():
        m = rh.build_roi_heads(cfg, {"res4": rs.ShapeSpec(channels=16, stride=16)}).train()
    gen = torch.Generator().manual_seed(31)
    d = dict(batch=np.int64(m.batch_size_per_image), pos_frac=np.float32(m.positive_sample_fraction),
             append_gt=np.int64(bool(m.proposal_append_gt)), num_classes=np.int64(K))
    props, targets = [], []
    for i, (n, n_obj, jit) in enumerate(((1500, 6, 0.3), (900, 3, 0.9), (700, 0, 0.3), (320, 2, 0.3))):
        h, w = 600, 800
        b, objs = synth_proposals(n, h, w, gen, n_obj=max(n_obj, 1))
        if jit > 0.5:          # most proposals hug the objects: more foreground than the cap
            k = int(0.8 * n)
            b[:k] = objs[torch.randint(0, len(objs), (k,), generator=gen)] + torch.randn(k, 4, generator=gen) * 3.0
            b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 1.0)
        p = rs.Instances((h, w))
        p.proposal_boxes = rs.Boxes(b)
        p.objectness_logits = torch.zeros(n)
        t = rs.Instances((h, w))
        t.gt_boxes = rs.Boxes(objs[:n_obj].clone())
        t.gt_classes = torch.randint(0, K, (n_obj,), generator=gen)
        props.append(p)
        targets.append(t)
        d["props%d" % i], d["gt_boxes%d" % i], d["gt_classes%d" % i] = b.numpy(), objs[:n_obj].numpy(), t.gt_classes.numpy()
    torch.manual_seed(7)
    out = m.label_and_sample_proposals(props, targets)
    for i, (o, p, t) in enumerate(zip(out, props, targets)):
        allp = torch.cat([p.proposal_boxes.tensor, t.gt_boxes.tensor], 0) if m.proposal_append_gt else p.proposal_boxes.tensor
        d["all_props%d" % i] = allp.numpy()
        if len(t) > 0:
            iou = rs.pairwise_iou(t.gt_boxes, rs.Boxes(allp))
            midx, mlab = m.proposal_matcher(iou)
            d["matched_idx%d" % i], d["matched_label%d" % i] = midx.numpy(), mlab.numpy().astype(np.int64)
        d["out_classes%d" % i] = o.gt_classes.numpy()
        d["out_props%d" % i] = o.proposal_boxes.tensor.numpy()
        d["out_gt%d" % i] = o.gt_boxes.tensor.numpy()
    np.savez(os.path.join(OUT, "label_sample.npz"), **d)


def gen_cosine():
    """Cosine similarity helpers of the reference (my_module.py:449-469: bsim_matrix, sim_matrix), run unchanged: the
    semantics behind the optional cosine + temperature logits against the text prototypes."""
    rs.install()
    mm = rs.load("defrcn.modeling.roi_heads.my_module")
    gen = torch.Generator().manual_seed(21)
    a = torch.relu(torch.randn(40, 64, generator=gen))
    a[3] = 0                                                   # zero row: the eps clamp
    t = torch.randn(21, 64, generator=gen)
    np.savez(os.path.join(OUT, "cosine.npz"), a=a.numpy(), t=t.numpy(), sim=mm.sim_matrix(a, t).numpy(),
             bsim=mm.bsim_matrix(a[None], t[None], tau=20.0)[0].numpy(), tau=np.float32(20.0))


def gen_rpn_select():
    """The reference's own find_top_rpn_proposals (proposal_generator/proposal_utils.py:13-118), eval mode, on
    synthetic RPN outputs: one level with ties / non-finite entries / boxes that clip to nothing, and three levels."""
    rs.install()
    pu = rs.load("defrcn.modeling.proposal_generator.proposal_utils")
    gen = torch.Generator().manual_seed(41)
    d = {}
    # no exact logit ties here: the reference asks torch for an UNSTABLE sort (:64), so their order is not defined by
    # it (torch 2.x's CPU sort really is unstable); the tie rule of the CUDA path (lower anchor first) is tested
    # against the oracle's stable sort instead
    cases = (("c4", 2, [3000], (600, 800), 0.7, 600, 200, 0.0, 0.0),
             ("c4_minsize", 2, [2000], (480, 672), 0.7, 2000, 1500, 4.0, 0.0),
             ("fpn3", 3, [1200, 300, 75], (600, 800), 0.6, 250, 300, 0.0, 0.0))
    for tag, N, sizes, (h, w), thr, pre, post, min_size, quant in cases:
        props, logits = synth_rpn_outputs(N, sizes, h, w, gen, quant)
        if tag == "c4":
            props[0][1, 17, 2] = float("inf")
            props[0][1, 400, 0] = float("nan")
            logits[0][1, 33] = float("nan")
            logits[0][0, 5] = float("inf")
            top = logits[0][0].argsort(descending=True)[:3]
            props[0][0, top[1]] = torch.tensor([900.0, 10.0, 950.0, 50.0])       # clips to zero width
            props[0][0, top[2]] = torch.tensor([50.0, 700.0, 90.0, 800.0])       # clips to zero height
        for lg in logits:
            for row in lg:
                fin = row[torch.isfinite(row)]
                assert len(torch.unique(fin)) == len(fin), "exact logit tie in a golden case"
        image_sizes = [(h, w) if i % 2 == 0 else (h - 24, w - 40) for i in range(N)]
        res = pu.find_top_rpn_proposals([p.clone() for p in props], [l.clone() for l in logits], image_sizes, thr, pre,
                                        post, min_size, False)
        d[tag + "_meta"] = np.array([N, len(sizes), pre, post], np.int64)
        d[tag + "_sizes"] = np.array(sizes, np.int64)
        d[tag + "_image_sizes"] = np.array(image_sizes, np.int64)
        d[tag + "_thr"] = np.float32(thr)
        d[tag + "_min_size"] = np.float32(min_size)
        for l, (p, lg) in enumerate(zip(props, logits)):
            d["%s_props%d" % (tag, l)], d["%s_logits%d" % (tag, l)] = p.numpy(), lg.numpy()
        for n, r in enumerate(res):
            d["%s_out_boxes%d" % (tag, n)] = r.proposal_boxes.tensor.numpy()
            d["%s_out_logits%d" % (tag, n)] = r.objectness_logits.numpy()
    np.savez_compressed(os.path.join(OUT, "rpn_select.npz"), **d)


def gen_eval_formats():
    """The reference's own PascalVOCDetectionEvaluator.process (pascal_voc_evaluation.py:57-78, with
    instances_to_coco_json :172-203) on synthetic detections of three images (one empty)."""
    rs.install()
    sys.modules.setdefault("defrcn.dataloader", types.ModuleType("defrcn.dataloader"))
    sys.modules["defrcn.dataloader"].build_detection_test_loader = None
    archs = types.ModuleType("defrcn.evaluation.archs")
    archs.resnet101 = None
    sys.modules["defrcn.evaluation.archs"] = archs

    class BoxMode:                      # detectron2 0.3 structures/boxes.py::BoxMode.convert, the one conversion used
        XYXY_ABS, XYWH_ABS = 0, 1

        @staticmethod
        def convert(box, from_mode, to_mode):
            assert (from_mode, to_mode) == (BoxMode.XYXY_ABS, BoxMode.XYWH_ABS)
            arr = torch.from_numpy(np.asarray(box)).clone()
            arr[:, 2] -= arr[:, 0]
            arr[:, 3] -= arr[:, 1]
            return arr.numpy()
    sys.modules["detectron2.structures"].BoxMode = BoxMode
    rs._mod("detectron2.utils.comm", is_main_process=lambda: True, gather=lambda x, dst=0: [x], synchronize=lambda: None)
    sys.modules["detectron2.utils"].comm = sys.modules["detectron2.utils.comm"]
    rs._mod("detectron2.utils.logger", create_small_table=lambda d: str(d))
    rs._mod("fvcore.common")
    rs._mod("fvcore.common.file_io", PathManager=object())
    pv = rs.load("defrcn.evaluation.pascal_voc_evaluation")
    ev = object.__new__(pv.PascalVOCDetectionEvaluator)
    ev._cpu_device = torch.device("cpu")
    ev.reset()
    gen = torch.Generator().manual_seed(17)
    ids = ["000012", "2008_000123", "000999"]
    counts = [37, 100, 0]
    T_ = 100
    boxes, scores, classes = np.zeros((3, T_, 4), np.float32), np.zeros((3, T_), np.float32), np.full((3, T_), -1, np.int64)
    inputs, outputs = [], []
    for i, (iid, n) in enumerate(zip(ids, counts)):
        b = torch.rand(n, 4, generator=gen) * 500
        b[:, 2:] = b[:, :2] + torch.rand(n, 2, generator=gen) * 333.3
        s = torch.rand(n, generator=gen).sort(descending=True).values
        if n > 4:
            s[1], s[2] = 0.99949997, 0.0005                     # rounding edges of :.3f
            b[0] = torch.tensor([0.04999, 10.25, 99.95, 100.05])
        c = torch.randint(0, 20, (n,), generator=gen)
        boxes[i, :n], scores[i, :n], classes[i, :n] = b.numpy(), s.numpy(), c.numpy()
        inst = rs.Instances((600, 800))
        inst.pred_boxes, inst.scores, inst.pred_classes = rs.Boxes(b.clone()), s.clone(), c.clone()
        inputs.append({"image_id": iid})
        outputs.append({"instances": inst})
    rs.Instances.to = lambda self, *a, **k: self
    ev.process(inputs, outputs)
    import json
    np.savez_compressed(os.path.join(OUT, "eval_formats.npz"), boxes=boxes, scores=scores, classes=classes,
                        counts=np.array(counts, np.int64), ids=np.array(ids),
                        voc_lines=np.array(json.dumps({str(k): v for k, v in ev._predictions.items()})),
                        coco=np.array(json.dumps([p["instances"] for p in ev._coco_preditions])))


def main():
    import sys
    if len(sys.argv) > 1:       # regenerate selected fixtures only: python -m oracle.gen_golden gen_train_step ...
        os.makedirs(OUT, exist_ok=True)
        torch.set_num_threads(1)
        for name in sys.argv[1:]:
            globals()[name]()
        return
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order in the CPU kernels
    gen_gdl()
    gen_fast_rcnn_inference()
    gen_losses()
    gen_attention()
    gen_head_tiny()
    gen_pcb()
    gen_teacher()
    gen_train_step()
    gen_train_step_cross()
    gen_known_answer()
    gen_cosine()
    gen_label_sample()
    gen_rpn_select()
    gen_eval_formats()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
