"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY.

CPU restatement ("port") of the reference's ROI-head hot path, used exclusively as the parity
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module; its CUDA path fails loudly without its extension.

What is restated, and where it lives in the reference (paths relative to /root/reference):
  roi_align_fwd/bwd        torchvision roi_align (third-party, torchvision==0.8.1, requirements.txt:77),
                           reached from defrcn/modeling/roi_heads/roi_heads.py:300-305,339-344
  apply_deltas / clip      detectron2==0.3 Box2BoxTransform / Boxes.clip, call sites fast_rcnn.py:306-324,108-110
  fast_rcnn_inference*     defrcn/modeling/roi_heads/fast_rcnn.py:46-134
  nms / batched_nms        torchvision nms + coordinate trick (fast_rcnn.py:125)
  gdl / affine             defrcn/modeling/meta_arch/gdl.py:6-38
  text_kv / siamese_attention / sematic_proposal_attention
                           defrcn/modeling/roi_heads/attentive_modules.py:36-55,58-75,114-177,262-294
  output_layers            defrcn/modeling/roi_heads/fast_rcnn.py:403-417 ; cross_output :462-476 + roi_heads.py:1154-1171
  pcb_calibrate            defrcn/evaluation/calibration_layer.py:106-124 (sklearn cosine_similarity semantics)
  res5 / head_forward      defrcn/modeling/roi_heads/roi_heads.py:313-344,1093-1149

Pinning status: the reference has NO tests or golden vectors of its own.  This port
is pinned against (a) outputs of the reference's own modules imported unchanged in the authoring
container (oracle/gen_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py) and
(b) the installed torchvision 0.26 CPU ops (tests/test_oracle_pinning.py).
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
SCALE_CLAMP = math.log(1000.0 / 16)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_nms.restype = ctypes.c_int64
        _LIB.oracle_batched_nms.restype = ctypes.c_int64
        _LIB.oracle_fast_rcnn_inference_single_image.restype = ctypes.c_int64
    return _LIB


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t,
                                dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ ROIAlign
def roi_align_fwd(feat, rois, pooled=(7, 7), spatial_scale=1.0 / 16, sampling_ratio=0, aligned=True,
                  impl="c"):
    """feat (N,C,H,W) fp32, rois (R,5) -> (R,C,PH,PW) torch fp32.  impl='c' is the restatement,
    impl='torchvision' the third-party op the reference actually calls."""
    if isinstance(pooled, int):
        pooled = (pooled, pooled)
    if impl == "torchvision":
        import torchvision
        return torchvision.ops.roi_align(torch.as_tensor(feat).float(), torch.as_tensor(rois).float(),
                                         pooled, spatial_scale, sampling_ratio, aligned)
    x, r = _f32(feat), _f32(rois)
    N, C, H, W = x.shape
    R = r.shape[0]
    out = np.zeros((R, C, pooled[0], pooled[1]), np.float32)
    rc = _lib().oracle_roi_align_fwd(_p(x), _p(r), N, C, H, W, R, pooled[0], pooled[1],
                                     ctypes.c_float(spatial_scale), int(sampling_ratio), int(aligned),
                                     _p(out))
    assert rc == 0
    return torch.from_numpy(out)


def roi_align_bwd(grad_out, rois, in_shape, spatial_scale=1.0 / 16, sampling_ratio=0, aligned=True):
    g, r = _f32(grad_out), _f32(rois)
    N, C, H, W = in_shape
    R, _, PH, PW = g.shape
    gin = np.zeros((N, C, H, W), np.float32)
    rc = _lib().oracle_roi_align_bwd(_p(g), _p(r), N, C, H, W, R, PH, PW, ctypes.c_float(spatial_scale),
                                     int(sampling_ratio), int(aligned), _p(gin))
    assert rc == 0
    return torch.from_numpy(gin)


def boxes_to_rois(box_lists):
    """detectron2 convert_boxes_to_pooler_format: list[(Ri,4)] -> (R,5) with batch index column."""
    return torch.cat([torch.cat([torch.full((len(b), 1), float(i)), torch.as_tensor(b).float()], dim=1)
                      for i, b in enumerate(box_lists)], dim=0)


# ------------------------------------------------------------------ decode / threshold / NMS
def apply_deltas(deltas, proposals, weights=(10.0, 10.0, 5.0, 5.0), impl="c"):
    """deltas (R,4K), proposals (R,4) -> (R,4K)."""
    if impl == "torch":
        d, b = torch.as_tensor(deltas).float(), torch.as_tensor(proposals).float()
        w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
        cx, cy = b[:, 0] + 0.5 * w, b[:, 1] + 0.5 * h
        dx, dy = d[:, 0::4] / weights[0], d[:, 1::4] / weights[1]
        dw = torch.clamp(d[:, 2::4] / weights[2], max=SCALE_CLAMP)
        dh = torch.clamp(d[:, 3::4] / weights[3], max=SCALE_CLAMP)
        pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
        pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
        out = torch.zeros_like(d)
        out[:, 0::4], out[:, 1::4] = pcx - 0.5 * pw, pcy - 0.5 * ph
        out[:, 2::4], out[:, 3::4] = pcx + 0.5 * pw, pcy + 0.5 * ph
        return out
    d, b = _f32(deltas), _f32(proposals)
    R, K = d.shape[0], d.shape[1] // 4
    out = np.zeros_like(d)
    _lib().oracle_apply_deltas(_p(d), _p(b), R, K, *(ctypes.c_float(w) for w in weights),
                               ctypes.c_float(SCALE_CLAMP), _p(out))
    return torch.from_numpy(out)


def nms(boxes, scores, thr):
    b, s = _f32(boxes), _f32(scores)
    n = b.shape[0]
    keep = np.zeros(max(n, 1), np.int64)
    nk = _lib().oracle_nms(_p(b), _p(s), ctypes.c_int64(n), ctypes.c_float(thr), _p(keep))
    return torch.from_numpy(keep[:nk].copy())


def batched_nms(boxes, scores, idxs, thr):
    b, s = _f32(boxes), _f32(scores)
    c = np.ascontiguousarray(torch.as_tensor(idxs).cpu().numpy(), dtype=np.int64)
    n = b.shape[0]
    keep = np.zeros(max(n, 1), np.int64)
    nk = _lib().oracle_batched_nms(_p(b), _p(s), _p(c), ctypes.c_int64(n), ctypes.c_float(thr), _p(keep))
    return torch.from_numpy(keep[:nk].copy())


def batched_nms_detectron2(boxes, scores, idxs, thr):
    """detectron2 v0.3 `layers.nms.batched_nms` as called from fast_rcnn.py:125: the coordinate-offset trick below 40 000
    boxes (`batched_nms` above), and from 40 000 boxes on per-class NMS on the un-offset boxes, survivors ordered by score
    (descending; the reference's argsort is unstable, so ties between survivors are only defined here: lower index first)."""
    n = int(torch.as_tensor(boxes).shape[0])
    if n < 40000:
        return batched_nms(boxes, scores, idxs, thr)
    b, s, c = torch.as_tensor(boxes).float(), torch.as_tensor(scores).float(), torch.as_tensor(idxs).long()
    kept = []
    for k in torch.unique(c).tolist():
        m = torch.nonzero(c == k).view(-1)
        kept.append(m[nms(b[m], s[m], thr)])
    keep = torch.sort(torch.cat(kept)).values
    return keep[torch.sort(s[keep], descending=True, stable=True).indices]


def fast_rcnn_inference_single_image(boxes, probs, image_shape, score_thresh=0.05, nms_thresh=0.5,
                                     topk=100):
    """boxes (R,4K) decoded/unclipped, probs (R,K+1).  Returns dict(boxes, scores, classes, roi_inds,
    n_candidates, cand_inds) following fast_rcnn.py:90-134."""
    b, p = _f32(boxes), _f32(probs)
    R, K = p.shape[0], p.shape[1] - 1
    cap = R * K if topk < 0 else min(R * K, topk)
    ob = np.zeros((max(cap, 1), 4), np.float32)
    os_ = np.zeros(max(cap, 1), np.float32)
    oc = np.zeros(max(cap, 1), np.int64)
    orr = np.zeros(max(cap, 1), np.int64)
    ncand = ctypes.c_int64(0)
    ci = np.zeros((max(R * K, 1), 2), np.int64)
    nk = _lib().oracle_fast_rcnn_inference_single_image(
        _p(b), _p(p), R, K, ctypes.c_float(image_shape[0]), ctypes.c_float(image_shape[1]),
        ctypes.c_float(score_thresh), ctypes.c_float(nms_thresh), ctypes.c_int64(topk), _p(ob), _p(os_),
        _p(oc), _p(orr), ctypes.byref(ncand), _p(ci))
    return dict(boxes=torch.from_numpy(ob[:nk].copy()), scores=torch.from_numpy(os_[:nk].copy()),
                classes=torch.from_numpy(oc[:nk].copy()), roi_inds=torch.from_numpy(orr[:nk].copy()),
                n_candidates=int(ncand.value), cand_inds=torch.from_numpy(ci[:ncand.value].copy()))


def fast_rcnn_inference(logits, deltas, proposals_per_image, image_shapes, score_thresh=0.05,
                        nms_thresh=0.5, topk=100, weights=(10.0, 10.0, 5.0, 5.0)):
    """FastRCNNOutputs.inference (fast_rcnn.py:306-360): softmax, decode, per-image post-processing."""
    probs = F.softmax(torch.as_tensor(logits).float(), dim=-1)
    props = torch.cat([torch.as_tensor(p).float() for p in proposals_per_image], dim=0)
    boxes = apply_deltas(deltas, props, weights, impl="torch")
    nper = [len(p) for p in proposals_per_image]
    return [fast_rcnn_inference_single_image(b, s, shp, score_thresh, nms_thresh, topk)
            for b, s, shp in zip(boxes.split(nper), probs.split(nper), image_shapes)]


# ------------------------------------------------------------------ GDL / affine
def affine_fwd(x, weight, bias=None):
    out = x * weight.expand_as(x)
    return out if bias is None else out + bias.expand_as(x)


def gdl_affine_bwd(grad_out, x, weight, lam):
    """d/dx of affine(decouple(x, lam)) and affine parameter grads."""
    gx = grad_out * weight.expand_as(grad_out) * lam
    gw = (grad_out * x).sum(dim=(0, 2, 3), keepdim=True)
    gb = grad_out.sum(dim=(0, 2, 3), keepdim=True)
    return gx, gw, gb


# ------------------------------------------------------------------ text fusion (fp32 torch)
def make_bg_embedding(class_embed, r):
    """create_normalized_orthogonal_tensor (utils/class_embedding.py:15-24) with explicit noise r."""
    mean = class_embed.mean(dim=0, keepdim=True)
    o = mean - torch.dot(mean.flatten(), r.flatten()) * r
    return o / torch.norm(o)


def text_kv(text_feat, p, prefix="attention."):
    """attentive_modules.py:274-277 then :125-126,129-135 — returns Kp (K+2,d), Vp (K+2,d)."""
    kt = F.relu(F.linear(text_feat, p[prefix + "key_projection.weight"], p[prefix + "key_projection.bias"]))
    vt = F.relu(F.linear(text_feat, p[prefix + "value_projection.weight"], p[prefix + "value_projection.bias"]))
    a = prefix + "attention."
    kp = F.linear(kt, p[a + "w_k.weight"])
    vp = F.linear(vt, p[a + "w_v.weight"])
    kp = torch.cat([kp, p[a + "dummy"].reshape(1, -1)], dim=0)
    vp = torch.cat([vp, torch.zeros(1, vp.shape[1], dtype=vp.dtype)], dim=0)
    return kp, vp


def siamese_attention(x, kp, vp, p, prefix="attention.attention."):
    """SingleHeadSiameseAttention.forward after the k/v projections (attentive_modules.py:123-177)."""
    d = x.shape[1]
    q = F.linear(x, p[prefix + "w_q.weight"])
    s = (q @ kp.t()) / np.power(d, 0.5)
    attn = F.softmax(s, dim=1)
    o = attn @ vp
    o1 = F.relu(F.linear(o * x, p[prefix + "linear1.0.weight"], p[prefix + "linear1.0.bias"]))
    o2 = F.relu(F.linear(x - o, p[prefix + "linear2.0.weight"], p[prefix + "linear2.0.bias"]))
    y = F.linear(torch.cat([o1, o2, x], dim=1), p[prefix + "linear3.weight"], p[prefix + "linear3.bias"])
    y2 = F.linear(F.relu(F.linear(y, p[prefix + "ffn.linear1.weight"], p[prefix + "ffn.linear1.bias"])),
                  p[prefix + "ffn.linear2.weight"], p[prefix + "ffn.linear2.bias"])
    z = F.layer_norm(y + y2, (d,), p[prefix + "ffn.norm3.weight"], p[prefix + "ffn.norm3.bias"], 1e-5)
    return z, attn


def sematic_proposal_attention(x, text_feat, p, prefix="attention."):
    """SematicProposalAttention.forward (attentive_modules.py:262-294): (sim2stext (R,d), attn (R,K+2))."""
    kp, vp = text_kv(text_feat, p, prefix)
    z, attn = siamese_attention(x, kp, vp, p, prefix + "attention.")
    return F.relu(z), attn


def output_layers(x, att_x, p, prefix="box_predictor."):
    """FastRCNNOutputLayers.forward in eval mode (fast_rcnn.py:403-417)."""
    deltas = F.linear(x, p[prefix + "bbox_pred.weight"], p[prefix + "bbox_pred.bias"])
    scores = F.linear(x if att_x is None else att_x, p[prefix + "cls_score.weight"], p[prefix + "cls_score.bias"])
    return scores, deltas


def cross_output_scores(sim2stext, text_feat, p):
    """SematicRes5ROIHeadsCrossOutput.forward_att (roi_heads.py:1157-1159): relu(out_proj(s)) @ T^T."""
    a = F.relu(F.linear(sim2stext, p["output_projection.weight"], p["output_projection.bias"]))
    return a @ text_feat.t()


def loss_attentive(attn, gt_classes):
    """roi_heads.py:1079-1081 — CE applied to attention *probabilities* (quirk preserved)."""
    return F.cross_entropy(attn, gt_classes, reduction="mean")


# ------------------------------------------------------------------ PCB
def pcb_calibrate(scores, feats, prototypes, classes, alpha=0.5, exclude=(), lower=0.05, upper=1.0):
    """calibration_layer.py:106-124.  scores (n,) sorted desc; feats (iright-ileft, D) for dets
    ileft..iright; prototypes (K,D).  sklearn cosine_similarity: normalise rows, then dot (fp32)."""
    s = torch.as_tensor(scores).float().clone()
    ileft = int((s > upper).sum())
    iright = int((s > lower).sum())
    f = torch.as_tensor(feats).float().numpy()
    pr = torch.as_tensor(prototypes).float().numpy()
    for i in range(ileft, iright):
        c = int(classes[i])
        if c in exclude:
            continue
        a = f[i - ileft]
        b = pr[c]
        na = np.sqrt(np.einsum("i,i->", a, a, dtype=np.float32)).astype(np.float32)
        nb = np.sqrt(np.einsum("i,i->", b, b, dtype=np.float32)).astype(np.float32)
        na = np.float32(1.0) if na == 0 else na
        nb = np.float32(1.0) if nb == 0 else nb
        cos = np.dot(a / na, b / nb)
        s[i] = s[i] * alpha + float(cos) * (1 - alpha)
    return s


# ------------------------------------------------------------------ res5 + whole head (CPU baseline)
def frozen_bn(x, p, pre):
    scale = p[pre + "weight"] * (p[pre + "running_var"] + 1e-5).rsqrt()
    bias = p[pre + "bias"] - p[pre + "running_mean"] * scale
    return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


def res5(x, p, prefix="res5.", stride_in_1x1=True):
    """roi_heads.py:313-344 — detectron2 make_stage(BottleneckBlock, 3, first_stride=2), FrozenBN."""
    for i in range(3):
        b = "%s%d." % (prefix, i)
        s = 2 if i == 0 else 1
        s1, s3 = (s, 1) if stride_in_1x1 else (1, s)
        out = F.relu(frozen_bn(F.conv2d(x, p[b + "conv1.weight"], stride=s1), p, b + "conv1.norm."))
        out = F.relu(frozen_bn(F.conv2d(out, p[b + "conv2.weight"], stride=s3, padding=1), p, b + "conv2.norm."))
        out = frozen_bn(F.conv2d(out, p[b + "conv3.weight"]), p, b + "conv3.norm.")
        sc = x
        if (b + "shortcut.weight") in p:
            sc = frozen_bn(F.conv2d(x, p[b + "shortcut.weight"], stride=s), p, b + "shortcut.norm.")
        x = F.relu(out + sc)
    return x


def head_forward(feat, proposals_per_image, image_shapes, text_feat, p, cross_output=False,
                 score_thresh=0.05, nms_thresh=0.5, topk=100, roi_impl="torchvision", stages=None):
    """SematicRes5ROIHeads.forward in eval mode (roi_heads.py:1093-1149) on CPU, fp32.
    `stages` (dict) receives per-stage wall times when given."""
    import time
    t = [time.perf_counter()]

    def tick(name):
        if stages is not None:
            now = time.perf_counter()
            stages[name] = stages.get(name, 0.0) + now - t[0]
            t[0] = now

    rois = boxes_to_rois(proposals_per_image)
    pooled = roi_align_fwd(feat, rois, 7, 1.0 / 16, 0, True, impl=roi_impl)
    tick("roi_align")
    x = res5(pooled, p).mean(dim=[2, 3])
    tick("res5")
    sim, attn = sematic_proposal_attention(x, text_feat, p)
    if cross_output:
        att_x = cross_output_scores(sim, text_feat, p)
        deltas = F.linear(x, p["box_predictor.bbox_pred.weight"], p["box_predictor.bbox_pred.bias"])
        logits = att_x
    else:
        logits, deltas = output_layers(x, sim, p)
    tick("text_fusion")
    dets = fast_rcnn_inference(logits, deltas, proposals_per_image, image_shapes, score_thresh,
                               nms_thresh, topk)
    tick("decode_nms")
    return dets, dict(pooled=pooled, feature_pooled=x, sim2stext=sim, attn=attn, logits=logits, deltas=deltas)


class _GDL(torch.autograd.Function):
    """meta_arch/gdl.py:6-16 — identity forward, grad * lambda backward."""

    @staticmethod
    def forward(ctx, x, lam):
        ctx.lam = lam
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.lam, None


def head_losses(logits, deltas, attn, gt_classes, proposals, gt_boxes, K, weights=(10.0, 10.0, 5.0, 5.0), beta=0.0):
    """FastRCNNOutputs.losses (fast_rcnn.py:222-304) + loss_attentive (roi_heads.py:1079-1081)."""
    from . import ref_stubs as rs
    loss_cls = F.cross_entropy(logits, gt_classes, reduction="mean")
    tgt = rs.Box2BoxTransform(weights).get_deltas(proposals, gt_boxes)
    fg = torch.nonzero((gt_classes >= 0) & (gt_classes < K)).squeeze(1)
    cols = 4 * gt_classes[fg][:, None] + torch.arange(4)
    n = (deltas[fg[:, None], cols] - tgt[fg]).abs()
    l1 = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    return {"loss_cls": loss_cls, "loss_box_reg": l1.sum() / gt_classes.numel(),
            "loss_attentive": loss_attentive(attn, gt_classes)}


def head_train_step(feat, proposals_per_image, gt_classes, gt_boxes, text_feat, p, aff_w, aff_b, lam, K, drop_p=0.0,
                    stages=None):
    """Fine-tune direction of the head on CPU, fp32 (rcnn.py:94-98 GDL + affine_rcnn, roi_heads.py:1093-1132 in train
    mode on pre-sampled, labelled proposals), differentiated by torch autograd: returns (losses, d(sum)/d(feat)); the
    parameter gradients are left in `.grad` of the tensors of `p` / aff_w / aff_b that require grad."""
    import time
    import torchvision
    t = [time.perf_counter()]

    def tick(name):
        if stages is not None:
            now = time.perf_counter()
            stages[name] = stages.get(name, 0.0) + now - t[0]
            t[0] = now

    feat = feat.detach().requires_grad_(True)
    f = _GDL.apply(feat, lam) * aff_w + aff_b
    rois = boxes_to_rois(proposals_per_image)
    pooled = torchvision.ops.roi_align(f, rois, (7, 7), 1.0 / 16, 0, True)
    tick("roi_align")
    x = res5(pooled, p).mean(dim=[2, 3])
    tick("res5")
    sim, attn = sematic_proposal_attention(x, text_feat, p)
    deltas = F.linear(x, p["box_predictor.bbox_pred.weight"], p["box_predictor.bbox_pred.bias"])
    logits = F.linear(F.dropout(sim, drop_p, training=drop_p > 0), p["box_predictor.cls_score.weight"],
                      p["box_predictor.cls_score.bias"])
    props = torch.cat([torch.as_tensor(b).float() for b in proposals_per_image], 0)
    losses = head_losses(logits, deltas, attn, gt_classes, props, gt_boxes, K)
    tick("text_fusion_losses")
    sum(losses.values()).backward()
    tick("backward")
    return losses, feat.grad


# ------------------------------------------------------------------ AP (parity of the end metric)
def voc_ap(rec, prec, use_07_metric=False):
    """defrcn/evaluation/pascal_voc_evaluation.py voc_ap (:227-258)."""
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def voc_eval_class(dets, gts, ovthresh=0.5, use_07_metric=False):
    """VOC AP for one class (pascal_voc_evaluation.py:261-372, `+1` pixel convention preserved).
    dets: list of (image_id, score, x1,y1,x2,y2); gts: {image_id: (M,4) array}."""
    npos = sum(len(v) for v in gts.values())
    seen = {k: np.zeros(len(v), bool) for k, v in gts.items()}
    if len(dets) == 0:
        return 0.0
    conf = np.array([d[1] for d in dets])
    order = np.argsort(-conf, kind="stable")
    tp, fp = np.zeros(len(dets)), np.zeros(len(dets))
    for di, j in enumerate(order):
        img, _, x1, y1, x2, y2 = dets[j]
        bb = np.array([x1, y1, x2, y2], float)
        g = np.asarray(gts.get(img, np.zeros((0, 4))), float)
        ovmax, jmax = -np.inf, -1
        if g.size > 0:
            ixmin, iymin = np.maximum(g[:, 0], bb[0]), np.maximum(g[:, 1], bb[1])
            ixmax, iymax = np.minimum(g[:, 2], bb[2]), np.minimum(g[:, 3], bb[3])
            iw, ih = np.maximum(ixmax - ixmin + 1.0, 0.0), np.maximum(iymax - iymin + 1.0, 0.0)
            inters = iw * ih
            uni = ((bb[2] - bb[0] + 1.0) * (bb[3] - bb[1] + 1.0)
                   + (g[:, 2] - g[:, 0] + 1.0) * (g[:, 3] - g[:, 1] + 1.0) - inters)
            ov = inters / uni
            ovmax, jmax = np.max(ov), int(np.argmax(ov))
        if ovmax > ovthresh and not seen[img][jmax]:
            tp[di] = 1.0
            seen[img][jmax] = True
        else:
            fp[di] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(max(npos, 1))
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return float(voc_ap(rec, prec, use_07_metric))


# ------------------------------------------------------------------ cosine logits (optional epilogue of C1')
def sim_matrix(a, b, eps=1e-12, tau=1.0):
    """my_module.py:461-469 `sim_matrix` (row-wise L2 normalisation with the norm clamped at eps, then a @ b^T), times a
    temperature as in `bsim_matrix` (:449-458)."""
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    a_n = a / torch.clamp(a.norm(dim=1)[:, None], min=eps)
    b_n = b / torch.clamp(b.norm(dim=1)[:, None], min=eps)
    return torch.mm(a_n, b_n.transpose(0, 1)) * tau


# ------------------------------------------------------------------ S1: proposal labelling / sampling
def label_proposals(props, gt_boxes, iou_thresh=0.5):
    """detectron2 pairwise_iou + Matcher(thresholds=[iou_thresh], labels=[0, 1], allow_low_quality_matches=False) as called
    from roi_heads.py:200-204: per proposal the ground-truth box of highest IoU (first maximum) and the fg (1) / bg (0)
    label.  Returns (matched_idx int64 (P,), matched_label int64 (P,), max_iou fp32 (P,))."""
    p, g = torch.as_tensor(props).float(), torch.as_tensor(gt_boxes).float().reshape(-1, 4)
    P = p.shape[0]
    if g.shape[0] == 0:
        return torch.zeros(P, dtype=torch.int64), torch.zeros(P, dtype=torch.int64), torch.zeros(P)
    a1 = (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1])
    a2 = (p[:, 2] - p[:, 0]) * (p[:, 3] - p[:, 1])
    wh = (torch.min(g[:, None, 2:], p[:, 2:]) - torch.max(g[:, None, :2], p[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    iou = torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(()))
    vals, idx = iou.max(dim=0)
    return idx, (vals >= iou_thresh).to(torch.int64), vals


def sample_counts(n_fg, n_bg, batch=512, positive_fraction=0.25):
    """detectron2 subsample_labels counts (roi_heads.py:118-155 `_sample_proposals`): (#fg rows, #bg rows)."""
    num_pos = min(n_fg, int(batch * positive_fraction))
    return num_pos, min(n_bg, batch - num_pos)


def label_and_sample(props, gt_boxes, gt_classes, K, batch=512, positive_fraction=0.25, iou_thresh=0.5, append_gt=True,
                     gen=None):
    """`ROIHeads.label_and_sample_proposals` for one image (roi_heads.py:157-250): ground truth appended to the proposals
    (PROPOSAL_APPEND_GT), argmax-IoU labelling, random subsample of <= batch * positive_fraction foreground and the
    rest background rows (torch.randperm, like detectron2's subsample_labels).  Returns (boxes, classes, matched gt boxes)."""
    p = torch.as_tensor(props).float()
    g = torch.as_tensor(gt_boxes).float().reshape(-1, 4)
    if append_gt:
        p = torch.cat([p, g], 0)
    idx, lab, _ = label_proposals(p, g, iou_thresh)
    cls = torch.as_tensor(gt_classes).to(torch.int64)[idx].clone() if g.shape[0] else torch.full((p.shape[0],), K, dtype=torch.int64)
    cls[lab == 0] = K
    pos, neg = torch.nonzero(cls != K).squeeze(1), torch.nonzero(cls == K).squeeze(1)
    n_pos, n_neg = sample_counts(pos.numel(), neg.numel(), batch, positive_fraction)
    sel = torch.cat([pos[torch.randperm(pos.numel(), generator=gen)[:n_pos]], neg[torch.randperm(neg.numel(), generator=gen)[:n_neg]]])
    gtb = g[idx[sel]] if g.shape[0] else torch.zeros((sel.numel(), 4))
    return p[sel], cls[sel], gtb


# ------------------------------------------------------------------ RPN proposal selection
def find_top_rpn_proposals(proposals, pred_objectness_logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_size=0.0):
    """detectron2 0.3 find_top_rpn_proposals as vendored at defrcn/modeling/proposal_generator/proposal_utils.py:13-118
    (eval behaviour: non-finite candidates are dropped).  proposals[l] (N,A_l,4), logits[l] (N,A_l).
    Returns per image dict(boxes, logits, n_invalid).  The descending sort is stable (what torch's CPU sort does;
    :64 asks for an unstable one, so ties are only pinned on the CPU)."""
    N = len(image_sizes)
    tb, ts, tl = [], [], []
    for lvl, (p, l) in enumerate(zip(proposals, pred_objectness_logits)):
        p, l = torch.as_tensor(p).float(), torch.as_tensor(l).float()
        k = min(pre_nms_topk, l.shape[1])
        sl, idx = l.sort(descending=True, dim=1, stable=True)                                  # :64
        idx = idx[:, :k]
        ts.append(sl[:, :k])                                                                   # :65
        tb.append(torch.gather(p, 1, idx[:, :, None].expand(-1, -1, 4)))                       # :69
        tl.append(torch.full((k,), lvl, dtype=torch.int64))
    ts, tb, tl = torch.cat(ts, 1), torch.cat(tb, 1), torch.cat(tl, 0)                          # :76-78
    out = []
    for n, (h, w) in enumerate(image_sizes):
        b, s, lv = tb[n].clone(), ts[n], tl
        valid = torch.isfinite(b).all(dim=1) & torch.isfinite(s)                               # :87
        n_invalid = int((~valid).sum())
        b, s, lv = b[valid], s[valid], lv[valid]
        b[:, 0::2] = b[:, 0::2].clamp(min=0, max=w)                                            # :96 Boxes.clip
        b[:, 1::2] = b[:, 1::2].clamp(min=0, max=h)
        keep = ((b[:, 2] - b[:, 0]) > min_box_size) & ((b[:, 3] - b[:, 1]) > min_box_size)     # :99 Boxes.nonempty
        b, s, lv = b[keep], s[keep], lv[keep]
        k = batched_nms(b, s, lv, nms_thresh)[:post_nms_topk]                                  # :103-111
        out.append(dict(boxes=b[k], logits=s[k], n_invalid=n_invalid))
    return out


# ------------------------------------------------------------------ detector_postprocess
def detector_postprocess(boxes, image_size, output_height, output_width):
    """detectron2 0.3 modeling/postprocessing.py::detector_postprocess on the boxes (call site
    defrcn/modeling/meta_arch/rcnn.py:69-73): scale by python-float ratios (torch multiplies the fp32 tensor by the
    ratio rounded to fp32), clip, keep non-empty.  Returns (boxes, keep_mask).  detectron2 is absent from
    /root/reference and from this image: this part is restated from the published v0.3 source, unpinned."""
    b = torch.as_tensor(boxes).float().clone()
    scale_x, scale_y = output_width / image_size[1], output_height / image_size[0]
    b[:, 0::2] *= scale_x
    b[:, 1::2] *= scale_y
    b[:, 0::2] = b[:, 0::2].clamp(min=0, max=output_width)
    b[:, 1::2] = b[:, 1::2].clamp(min=0, max=output_height)
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    return b[keep], keep
