"""Fast R-CNN outputs for the B200 ROI head.

API mirror of defrcn/modeling/roi_heads/fast_rcnn.py (same registry, class and function names):
`fast_rcnn_inference(_single_image)` (:46-134), `FastRCNNOutputs` (:137-360), `FastRCNNOutputLayers`
(:363-417), `FastRCNNAttentionOutputLayers` (:422-476).  Inference post-processing (softmax, decode, clip,
threshold, per-class NMS, top-k) runs on the device through csrc/detect_post.cu with a single
device->host read (the per-image detection counts) where the reference synchronises per image.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from ... import ops
from ...layers import cat, get_event_storage, smooth_l1_loss
from ...structures import Boxes, Instances, Registry

ROI_HEADS_OUTPUT_REGISTRY = Registry("ROI_HEADS_OUTPUT")


def _instances_from_padded(out, image_shapes):
    counts = out["counts"].tolist()          # the only host synchronisation of the post-processing
    results, kept = [], []
    for i, (shape, n) in enumerate(zip(image_shapes, counts)):
        inst = Instances(shape)
        inst.pred_boxes = Boxes(out["boxes"][i, :n])
        inst.scores = out["scores"][i, :n]
        inst.pred_classes = out["classes"][i, :n]
        results.append(inst)
        kept.append(out["roi_inds"][i, :n])
    return results, kept


def fast_rcnn_inference(boxes, scores, image_shapes, score_thresh, nms_thresh, topk_per_image):
    """boxes: list of (Ri, 4K) decoded (unclipped) boxes, scores: list of (Ri, K+1) probabilities."""
    res = [fast_rcnn_inference_single_image(b, s, shp, score_thresh, nms_thresh, topk_per_image)
           for s, b, shp in zip(scores, boxes, image_shapes)]
    return tuple(list(x) for x in zip(*res))


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """Single-image entry with the reference's signature (fast_rcnn.py:90-134): `boxes` (R, 4K) or (R, 4) are
    already decoded, `scores` (R, K+1) are probabilities.  Clip + threshold use torch indexing (same row-major
    order as `nonzero`), the per-class NMS and top-k run in csrc/detect_post.cu.  The head itself does not come
    through here: it uses the fused logits -> detections path (`FastRCNNOutputs.inference`)."""
    dev = scores.device
    R, K = scores.shape[0], scores.shape[1] - 1
    nb = boxes.shape[1] // 4
    h, w = float(image_shape[0]), float(image_shape[1])
    b = boxes.reshape(R, nb, 4).float().clone()
    b[..., 0::2].clamp_(min=0, max=w)
    b[..., 1::2].clamp_(min=0, max=h)
    fg = scores[:, :K].float()
    mask = fg > score_thresh
    idx = mask.nonzero()
    cb = (b[idx[:, 0], 0] if nb == 1 else b[mask]).contiguous()
    cs = fg[mask].contiguous()
    n = cb.shape[0]
    result = Instances(image_shape)
    if n == 0:
        result.pred_boxes = Boxes(cb)
        result.scores = cs
        result.pred_classes = idx[:, 1]
        return result, idx[:, 0]
    topk = n if topk_per_image < 0 else min(n, topk_per_image)
    so = torch.zeros(1, dtype=torch.int32, device=dev)
    sc = torch.full((1,), n, dtype=torch.int32, device=dev)
    keep, kc = ops.batched_nms_segments(cb, cs, idx[:, 1].to(torch.int32).contiguous(), so, sc, K, nms_thresh, topk)
    k = keep[0, :int(kc.item())].to(torch.int64)
    result.pred_boxes = Boxes(cb[k])
    result.scores = cs[k]
    result.pred_classes = idx[k, 1]
    return result, idx[k, 0]


class FastRCNNOutputs(object):
    """Holds the head outputs of one batch; `losses()` for training, `inference()` for testing."""

    def __init__(self, box2box_transform, pred_class_logits, pred_proposal_deltas, proposals, smooth_l1_beta,
                 Guided_gt_classes=None):
        self.box2box_transform = box2box_transform
        self.num_preds_per_image = [len(p) for p in proposals]
        self.pred_class_logits = pred_class_logits
        self.pred_proposal_deltas = pred_proposal_deltas
        self.smooth_l1_beta = smooth_l1_beta
        self.proposals = Boxes.cat([p.proposal_boxes for p in proposals])
        assert not self.proposals.tensor.requires_grad, "Proposals should not require gradients!"
        self.image_shapes = [x.image_size for x in proposals]
        if proposals[0].has("gt_boxes"):
            self.gt_boxes = Boxes.cat([p.gt_boxes for p in proposals])
            assert proposals[0].has("gt_classes")
            self.gt_classes = cat([p.gt_classes for p in proposals], dim=0)
        self.Guided_gt_classes = Guided_gt_classes

    # ---- training ---------------------------------------------------------------------------------
    def _log_accuracy(self, defer=None):
        """fast_rcnn.py:191-220.  defer(key, tensor): hand the counts over without a host read (static-shape masks instead
        of boolean indexing, asynchronous copy; `ROIHeads.flush_deferred_logs` logs them one step later)."""
        n = self.gt_classes.numel()
        pred = self.pred_class_logits.argmax(dim=1)
        bg = self.pred_class_logits.shape[1] - 1
        fg = (self.gt_classes >= 0) & (self.gt_classes < bg)
        if defer is not None:
            hit = pred == self.gt_classes
            defer("accuracy", torch.stack([hit.sum(), fg.sum(), (hit & fg).sum(), ((pred == bg) & fg).sum(),
                                           torch.full_like(fg.sum(), n)]))
            return
        stats = torch.stack([(pred == self.gt_classes).sum(), fg.sum(), (pred[fg] == self.gt_classes[fg]).sum(),
                             (pred[fg] == bg).sum()]).tolist()   # one sync instead of four
        acc, n_fg, fg_acc, fn = stats
        st = get_event_storage()
        st.put_scalar("fast_rcnn/cls_accuracy", acc / max(n, 1))
        if n_fg > 0:
            st.put_scalar("fast_rcnn/fg_cls_accuracy", fg_acc / n_fg)
            st.put_scalar("fast_rcnn/false_negative", fn / n_fg)

    def softmax_cross_entropy_loss(self):
        self._log_accuracy()
        return F.cross_entropy(self.pred_class_logits, self.gt_classes, reduction="mean")

    def smooth_l1_loss(self):
        target = self.box2box_transform.get_deltas(self.proposals.tensor, self.gt_boxes.tensor)
        box_dim = target.size(1)
        agnostic = self.pred_proposal_deltas.size(1) == box_dim
        bg = self.pred_class_logits.shape[1] - 1
        fg = torch.nonzero((self.gt_classes >= 0) & (self.gt_classes < bg)).squeeze(1)
        ar = torch.arange(box_dim, device=self.pred_proposal_deltas.device)
        cols = ar if agnostic else box_dim * self.gt_classes[fg][:, None] + ar
        loss = smooth_l1_loss(self.pred_proposal_deltas[fg[:, None], cols], target[fg], self.smooth_l1_beta, reduction="sum")
        return loss / self.gt_classes.numel()    # normalised by R, not by the number of foreground rows

    def losses(self):
        return {"loss_cls": self.softmax_cross_entropy_loss(), "loss_box_reg": self.smooth_l1_loss()}

    # ---- inference --------------------------------------------------------------------------------
    def predict_boxes(self):
        R, B = len(self.proposals), self.proposals.tensor.shape[1]
        K = self.pred_proposal_deltas.shape[1] // B
        boxes = self.box2box_transform.apply_deltas(
            self.pred_proposal_deltas.view(R * K, B), self.proposals.tensor.unsqueeze(1).expand(R, K, B).reshape(-1, B))
        return boxes.view(R, K * B).split(self.num_preds_per_image, dim=0)

    def predict_probs(self):
        return F.softmax(self.pred_class_logits, dim=-1).split(self.num_preds_per_image, dim=0)

    def inference_device(self, score_thresh, nms_thresh, topk_per_image):
        """Fused logits -> padded detections, no host sync (see ops.fast_rcnn_inference_device)."""
        dev = self.pred_class_logits.device
        _, offs = ops._roi_index(tuple(int(n) for n in self.num_preds_per_image), dev)     # cached: no H2D per step
        hw = ops.image_hw_tensor(self.image_shapes, dev)
        return ops.fast_rcnn_inference_device(self.pred_class_logits, self.pred_proposal_deltas, self.proposals.tensor,
                                              offs, hw, score_thresh, nms_thresh, topk_per_image,
                                              weights=self.box2box_transform.weights,
                                              max_rois_per_image=max(self.num_preds_per_image) if self.num_preds_per_image else None)

    def inference(self, score_thresh, nms_thresh, topk_per_image):
        out = self.inference_device(score_thresh, nms_thresh, topk_per_image)
        return _instances_from_padded(out, self.image_shapes)


class _OutputLayersBase(nn.Module):
    def __init__(self, cfg, input_size, num_classes, cls_agnostic_bbox_reg, box_dim=4):
        super().__init__()
        if not isinstance(input_size, int):
            input_size = int(np.prod(input_size))
        self.cls_score = nn.Linear(input_size, num_classes + 1)
        self.bbox_pred = nn.Linear(input_size, (1 if cls_agnostic_bbox_reg else num_classes) * box_dim)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        nn.init.constant_(self.cls_score.bias, 0)
        nn.init.constant_(self.bbox_pred.bias, 0)
        self._do_cls_dropout = cfg.MODEL.ROI_HEADS.CLS_DROPOUT
        self._dropout_ratio = cfg.MODEL.ROI_HEADS.DROPOUT_RATIO
        self._w = {}

    def _bf16(self, name, p):
        key = (p.data_ptr(), p._version, ops.PARAM_GENERATION[0])
        hit = self._w.get(name)
        if hit is None or hit[0] != key:
            hit = (key, p.detach().to(torch.bfloat16).contiguous())
            self._w[name] = hit
        return hit[1]

    def _linear(self, layer, name, x, x_bf16=None):
        """bf16 tcgen05 GEMM at inference on CUDA; autograd path (library GEMM) while training."""
        if self.training and torch.is_grad_enabled():
            return layer(x.float())
        if x_bf16 is None:
            x_bf16 = x.to(torch.bfloat16) if x.dtype != torch.bfloat16 else x
        return ops.gemm_bf16(x_bf16, self._bf16(name, layer.weight), layer.bias)


@ROI_HEADS_OUTPUT_REGISTRY.register()
class FastRCNNOutputLayers(_OutputLayersBase):
    """bbox_pred on the visual feature, cls_score on the (text-fused) feature (fast_rcnn.py:403-417)."""

    def forward(self, x, att_x=None, x_bf16=None, att_x_bf16=None):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        proposal_deltas = self._linear(self.bbox_pred, "bbox_pred", x, x_bf16)
        if att_x is None:
            att_x, att_x_bf16 = x, x_bf16
        if self._do_cls_dropout and self.training:
            att_x, att_x_bf16 = F.dropout(att_x, self._dropout_ratio, training=True), None
        scores = self._linear(self.cls_score, "cls_score", att_x, att_x_bf16)
        return scores, proposal_deltas


@ROI_HEADS_OUTPUT_REGISTRY.register()
class FastRCNNAttentionOutputLayers(_OutputLayersBase):
    """`att_x` already holds the class logits (dot products against the text prototypes); only bbox_pred is
    applied here (fast_rcnn.py:462-476).  cls_score exists for checkpoint compatibility and is unused."""

    def forward(self, x, att_x=None, x_bf16=None, att_x_bf16=None):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        proposal_deltas = self._linear(self.bbox_pred, "bbox_pred", x, x_bf16)
        att_x = x if att_x is None else att_x
        if self._do_cls_dropout and self.training:
            att_x = F.dropout(att_x, self._dropout_ratio, training=True)
        return att_x, proposal_deltas
