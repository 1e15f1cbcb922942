"""Oracle port vs golden vectors produced by the reference's own modules (oracle/gen_golden.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def params(g, strip=""):
    return {k[len(strip):] if strip and k.startswith(strip) else k: T(g[k]) for k in g.files}


def test_gdl_affine(golden):
    g = golden("gdl")
    y = O.affine_fwd(T(g["x"]), T(g["w"]), T(g["b"]))
    assert torch.equal(y, T(g["y"]))
    gx, gw, gb = O.gdl_affine_bwd(T(g["g"]), T(g["x"]), T(g["w"]), float(g["lam"]))
    assert torch.allclose(gx, T(g["gx"]), rtol=1e-6, atol=1e-9)
    assert torch.allclose(gw, T(g["gw"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(gb, T(g["gb"]), rtol=1e-5, atol=1e-6)


def test_fast_rcnn_inference(golden):
    g = golden("fast_rcnn_inference")
    for tag in ("voc", "coco", "ties", "empty"):
        h, w = g[tag + "_hw"]
        # decode: C restatement and torch restatement vs reference Box2BoxTransform
        for impl in ("c", "torch"):
            b = O.apply_deltas(g[tag + "_deltas"], g[tag + "_props"], impl=impl)
            assert torch.allclose(b, T(g[tag + "_pred_boxes_all"]), rtol=1e-5, atol=1e-3), (tag, impl)
        # post-processing is bit-exact given the reference's own decoded boxes / probabilities
        r = O.fast_rcnn_inference_single_image(g[tag + "_pred_boxes_all"], g[tag + "_probs"], (h, w), 0.05, 0.5, 100)
        assert r["n_candidates"] == int(g[tag + "_ncand"]), tag
        assert torch.equal(r["classes"], T(g[tag + "_classes"])), tag
        assert torch.equal(r["roi_inds"], T(g[tag + "_roi_inds"])), tag
        assert torch.equal(r["scores"], T(g[tag + "_scores"])), tag
        assert torch.equal(r["boxes"], T(g[tag + "_boxes"])), tag
    assert int(g["empty_ncand"]) == 0 and len(g["empty_scores"]) == 0
    assert len(g["ties_scores"]) > 0


def test_losses(golden):
    g = golden("losses")
    from oracle import ref_stubs as rs
    K = g["logits"].shape[1] - 1
    loss_cls = F.cross_entropy(T(g["logits"]), T(g["gt_classes"]))
    assert abs(float(loss_cls) - float(g["loss_cls"])) < 1e-6
    tgt = rs.Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).get_deltas(T(g["props"]), T(g["gt_boxes"]))
    fg = torch.nonzero((T(g["gt_classes"]) >= 0) & (T(g["gt_classes"]) < K)).squeeze(1)
    cols = 4 * T(g["gt_classes"])[fg][:, None] + torch.arange(4)
    l1 = (T(g["deltas"])[fg[:, None], cols] - tgt[fg]).abs().sum() / len(g["gt_classes"])
    assert abs(float(l1) - float(g["loss_box_reg"])) < 1e-5


def test_attention_small(golden):
    g = golden("attention_small")
    p = params(g)
    text = torch.cat([T(g["embed"]), T(g["bg_feature"])], 0)
    assert torch.equal(text, T(g["text_feat"]))
    sim, attn = O.sematic_proposal_attention(T(g["x"]), text, p)
    assert attn.shape == (g["x"].shape[0], text.shape[0] + 1)
    assert torch.allclose(attn, T(g["attn"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(sim, T(g["sim2stext"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(attn.sum(1), torch.ones(attn.shape[0]), atol=1e-5)


def _head(golden, tag, cross):
    g = golden(tag)
    p = params(g)
    text = torch.cat([T(g["embed"]), T(g["bg_feature"])], 0)
    props = [T(g["props0"]), T(g["props1"])]
    shapes = [tuple(g["hw0"]), tuple(g["hw1"])]
    for impl in ("c", "torchvision"):
        dets, mid = O.head_forward(T(g["feat"]), props, shapes, text, p, cross_output=cross, roi_impl=impl)
        assert torch.allclose(mid["pooled"], T(g["pooled"]), rtol=1e-5, atol=1e-6)
        assert torch.allclose(mid["feature_pooled"], T(g["feature_pooled"]), rtol=1e-4, atol=1e-5)
        assert torch.allclose(mid["sim2stext"], T(g["sim2stext"]), rtol=1e-3, atol=1e-4)
        assert torch.allclose(mid["logits"], T(g["logits"]), rtol=1e-3, atol=1e-4)
        assert torch.allclose(mid["deltas"], T(g["deltas"]), rtol=1e-3, atol=1e-4)
        for i, d in enumerate(dets):
            assert len(g["det_scores%d" % i]) > 0
            assert torch.equal(d["classes"], T(g["det_classes%d" % i]))
            assert torch.allclose(d["scores"], T(g["det_scores%d" % i]), rtol=1e-4, atol=1e-6)
            assert torch.allclose(d["boxes"], T(g["det_boxes%d" % i]), rtol=1e-4, atol=1e-2)


def test_head_tiny(golden):
    _head(golden, "head_tiny", False)


def test_head_tiny_cross(golden):
    _head(golden, "head_tiny_cross", True)


def test_pcb(golden):
    g = golden("pcb")
    # Q1: ROIAlign 1x1 @ 1/32 adaptive + fc
    rois = O.boxes_to_rois([T(g["boxes"])])
    pooled = O.roi_align_fwd(g["conv_feature"], rois, 1, 1.0 / 32, 0, True)
    feats = F.linear(pooled.flatten(1), T(g["fc_w"]), T(g["fc_b"]))
    assert torch.allclose(feats, T(g["feats_all"]), rtol=1e-4, atol=1e-5)
    # Q2: calibration over ileft..iright
    s_in = T(g["scores_in"])
    ileft, iright = int((s_in > 1.0).sum()), int((s_in > 0.05).sum())
    assert 0 < ileft < iright < len(s_in)
    out = O.pcb_calibrate(s_in, T(g["feats_all"])[ileft:iright], T(g["protos"]), g["classes"], 0.5,
                          set(g["exclude"].tolist()))
    assert torch.allclose(out, T(g["scores_out"]), rtol=1e-5, atol=1e-6)
    assert not torch.equal(out, s_in)


def test_known_answer_ce(golden):
    g = golden("known_answer_ce")
    assert abs(float(g["ce"]) - 2.764926910) < 1e-6          # SURVEY.md §4 / BASELINE.md
    assert tuple(g["shape"]) == (2048, 16) and int(g["n_bg"]) == 1697


def test_train_step_autograd(golden):
    """The oracle restatement, differentiated by torch autograd on the CPU, reproduces the reference's own losses and
    gradients for the fine-tune direction (fixture from SematicRes5ROIHeads.forward_att + FastRCNNOutputs.losses)."""
    g = golden("train_step")
    from oracle import ref_stubs as rs
    K = 20
    p = {k: T(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith(("attention.", "box_predictor."))}
    x = T(g["x"]).clone().requires_grad_(True)
    text = torch.cat([T(g["embed"]), T(g["bg_feature"])], 0)
    sim, attn = O.sematic_proposal_attention(x, text, p)
    logits, deltas = O.output_layers(x, sim, p)
    gt = T(g["gt_classes"])
    loss_cls = F.cross_entropy(logits, gt)
    tgt = rs.Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).get_deltas(T(g["props"]), T(g["gt_boxes"]))
    fg = torch.nonzero((gt >= 0) & (gt < K)).squeeze(1)
    cols = 4 * gt[fg][:, None] + torch.arange(4)
    loss_box = (deltas[fg[:, None], cols] - tgt[fg]).abs().sum() / gt.numel()
    loss_att = O.loss_attentive(attn, gt)
    for name, v in (("loss_cls", loss_cls), ("loss_box_reg", loss_box), ("loss_attentive", loss_att)):
        assert abs(float(v) - float(g["loss." + name])) < 1e-4 * max(1.0, abs(float(g["loss." + name]))), name
    (loss_cls + loss_box + loss_att).backward()
    assert torch.allclose(x.grad, T(g["grad_x"]), rtol=1e-3, atol=1e-5)
    n = 0
    for k, v in p.items():
        if "grad." + k in g.files:
            ref = T(g["grad." + k])
            assert float((v.grad - ref).norm()) <= 1e-3 * float(ref.norm()) + 1e-6, k
            n += 1
    assert n >= 24


def test_cosine_similarity_helpers(golden):
    """Oracle restatement of the reference's cosine helpers (my_module.py:449-469) against their own outputs — the
    semantics of the optional cosine + temperature logits (zero row: the eps clamp)."""
    g = golden("cosine")
    assert torch.allclose(O.sim_matrix(g["a"], g["t"]), T(g["sim"]), rtol=1e-6, atol=1e-7)
    # bsim_matrix normalises with F.normalize (same eps) and multiplies by tau
    assert torch.allclose(O.sim_matrix(g["a"], g["t"], tau=float(g["tau"])), T(g["bsim"]), rtol=1e-5, atol=1e-5)
    assert float(O.sim_matrix(g["a"], g["t"])[3].abs().max()) == 0.0


def test_label_proposals_vs_reference(golden):
    """Oracle restatement of the matching / sampling counts of ROIHeads.label_and_sample_proposals against the
    reference's own run (fixture: 3 images with ground truth — few objects, more foreground than the cap, fewer
    proposals than the batch — and one without)."""
    g = golden("label_sample")
    B, frac, K = int(g["batch"]), float(g["pos_frac"]), int(g["num_classes"])
    for i in (0, 1, 3):
        idx, lab, _ = O.label_proposals(g["all_props%d" % i], g["gt_boxes%d" % i])
        assert torch.equal(idx, T(g["matched_idx%d" % i])) and torch.equal(lab, T(g["matched_label%d" % i]))
        n_fg = int(lab.sum())
        npos, nneg = O.sample_counts(n_fg, len(lab) - n_fg, B, frac)
        oc = T(g["out_classes%d" % i])
        assert (int((oc < K).sum()), int((oc == K).sum())) == (npos, nneg)
    idx, lab, _ = O.label_proposals(g["all_props2"], g["gt_boxes2"])
    assert int(lab.sum()) == 0 and O.sample_counts(0, len(lab), B, frac) == (0, len(g["out_classes2"]))


def _rpn_case(g, tag):
    N, L, pre, post = (int(v) for v in g[tag + "_meta"])
    props = [T(g["%s_props%d" % (tag, l)]) for l in range(L)]
    logits = [T(g["%s_logits%d" % (tag, l)]) for l in range(L)]
    image_sizes = [tuple(int(v) for v in s) for s in g[tag + "_image_sizes"]]
    return N, props, logits, image_sizes, float(g[tag + "_thr"]), pre, post, float(g[tag + "_min_size"])


def test_rpn_select(golden):
    """find_top_rpn_proposals restatement vs the reference's own function (proposal_utils.py:13-118): bit-exact."""
    g = golden("rpn_select")
    for tag in ("c4", "c4_minsize", "fpn3"):
        N, props, logits, image_sizes, thr, pre, post, min_size = _rpn_case(g, tag)
        out = O.find_top_rpn_proposals(props, logits, image_sizes, thr, pre, post, min_size)
        for n in range(N):
            assert torch.equal(out[n]["boxes"], T(g["%s_out_boxes%d" % (tag, n)])), (tag, n)
            assert torch.equal(out[n]["logits"], T(g["%s_out_logits%d" % (tag, n)])), (tag, n)
    assert O.find_top_rpn_proposals(*_rpn_case(g, "c4")[1:])[1]["n_invalid"] >= 1     # the planted inf / NaN entries


def test_detector_postprocess_restatement():
    """detector_postprocess restatement against the Boxes.scale / clip / nonempty stubs used for the reference run."""
    from oracle import ref_stubs as rs
    gen = torch.Generator().manual_seed(3)
    b = torch.rand(64, 4, generator=gen) * 600
    b[:, 2:] = b[:, :2] + torch.rand(64, 2, generator=gen) * 300
    b[5] = torch.tensor([700.0, 10.0, 790.0, 40.0])
    h, w, oh, ow = 600, 800, 375, 500
    out, keep = O.detector_postprocess(b, (h, w), oh, ow)
    bx = rs.Boxes(b.clone())
    bx.scale(ow / w, oh / h)
    bx.clip((oh, ow))
    k = bx.nonempty()
    assert torch.equal(keep, k) and torch.equal(out, bx.tensor[k])
