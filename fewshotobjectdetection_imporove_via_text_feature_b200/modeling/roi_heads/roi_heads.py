"""ROI heads, B200-native, behind Detectron2's ROI_HEADS_REGISTRY / `forward(images, features, proposals, targets)`.

API mirror of defrcn/modeling/roi_heads/roi_heads.py: `ROI_HEADS_REGISTRY`, `build_roi_heads` (:27-45),
`ROIHeads` (:78-277), `Res5ROIHeads` (:280-386), `SematicRes5ROIHeads` (:921-1149),
`SematicRes5ROIHeadsCrossOutput` (:1150-1171).  State-dict names match (`res5.*`, `box_predictor.*`,
`attention.*`, `output_projection.*`, `sematic_projection.*`, `projection_matrix`).

Hot path in eval mode: ROIAlign kernel (channels-last gather) -> res5 (cuDNN, FrozenBN folded, bf16 channels-last)
-> spatial mean -> tcgen05 text-fusion chain -> predictor GEMMs -> fused softmax/decode/threshold/NMS kernels.
"""
import logging
import os
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from ...config import b200_opt
from ...layers import (BottleneckBlock, Box2BoxTransform, Matcher, add_ground_truth_to_proposals, cat, frozen_res5_mean,
                       get_event_storage, make_stage, nonzero_tuple, subsample_labels)
from ...structures import Boxes, Instances, Registry, ShapeSpec, pairwise_iou
from ..poolers import ROIPooler
from .attentive_modules import SEMANTIC_DIM, SematicProposalAttention
from .fast_rcnn import ROI_HEADS_OUTPUT_REGISTRY, FastRCNNOutputs

ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
logger = logging.getLogger(__name__)


def build_roi_heads(cfg, input_shape):
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)


def select_foreground_proposals(proposals, bg_label):
    fg, masks = [], []
    for p in proposals:
        m = (p.gt_classes != -1) & (p.gt_classes != bg_label)
        fg.append(p[m.nonzero().squeeze(1)])
        masks.append(m)
    return fg, masks


class ROIHeads(nn.Module):
    """Per-region computation base: proposal matching / sampling shared by all heads."""

    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        rh, bh = cfg.MODEL.ROI_HEADS, cfg.MODEL.ROI_BOX_HEAD
        self.batch_size_per_image = rh.BATCH_SIZE_PER_IMAGE
        self.positive_sample_fraction = rh.POSITIVE_FRACTION
        self.test_score_thresh = rh.SCORE_THRESH_TEST
        self.test_nms_thresh = rh.NMS_THRESH_TEST
        self.test_detections_per_img = cfg.TEST.DETECTIONS_PER_IMAGE
        self.in_features = rh.IN_FEATURES
        self.num_classes = rh.NUM_CLASSES
        self.proposal_append_gt = rh.PROPOSAL_APPEND_GT
        self.feature_strides = {k: v.stride for k, v in input_shape.items()}
        self.feature_channels = {k: v.channels for k, v in input_shape.items()}
        self.cls_agnostic_bbox_reg = bh.CLS_AGNOSTIC_BBOX_REG
        self.smooth_l1_beta = bh.SMOOTH_L1_BETA
        self.proposal_matcher = Matcher(rh.IOU_THRESHOLDS, rh.IOU_LABELS, allow_low_quality_matches=False)
        self.box2box_transform = Box2BoxTransform(weights=bh.BBOX_REG_WEIGHTS)
        # STATIC_SAMPLING: the training forward makes no host read at all (CUDA-graph capturable): every image contributes
        # exactly BATCH_SIZE_PER_IMAGE sampled rows (true whenever an image has that many background candidates, i.e.
        # always with RPN's 2000 proposals; a device flag records a violation), and the logging scalars the reference
        # reads back synchronously (roi_heads.py:236-248, fast_rcnn.py:191-220) are copied to pinned host memory
        # asynchronously and handed to the event storage by `flush_deferred_logs()` one step later.
        self.static_sampling = bool(b200_opt(cfg, "STATIC_SAMPLING", False))
        self._deferred = {}

    def _mark(self, name, tensor=None):
        """Instrumentation hook (bench.py sets `_stage_cb`): called at the stage boundaries inside `forward` with the
        tensor the stage produced, so that a caller can record CUDA events / gradient hooks without reaching into the
        forward's internals.  No-op otherwise."""
        cb = getattr(self, "_stage_cb", None)
        if cb is not None:
            cb(name, tensor)

    def _assign_labels(self, matched_idxs, matched_labels, gt_classes):
        if gt_classes.numel() > 0:
            gt_classes = gt_classes[matched_idxs]
            gt_classes[matched_labels == 0] = self.num_classes   # unmatched -> background
            gt_classes[matched_labels == -1] = -1                # ignore
        else:
            gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
        return gt_classes

    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes):
        gt_classes = self._assign_labels(matched_idxs, matched_labels, gt_classes)
        fg, bg = subsample_labels(gt_classes, self.batch_size_per_image, self.positive_sample_fraction, self.num_classes)
        idx = torch.cat([fg, bg], dim=0)
        return idx, gt_classes[idx]

    def _label_and_sample_device(self, proposals, targets, append_gt=False):
        """One kernel for the whole batch (ops.label_and_sample_proposals) and one small device->host read (the per-image
        row counts, which the reference's logging needs on the host anyway) instead of a Python loop of torch ops with
        several synchronisations per image.  Labels / matches are the reference's bit for bit; the random subsample
        has its distribution but not torch's RNG stream.  append_gt: the kernel takes the ground-truth boxes as extra
        candidates itself (static sampling only; otherwise the caller appended them to `proposals`)."""
        r = ops.label_and_sample_proposals([p.proposal_boxes.tensor for p in proposals], [t.gt_boxes.tensor for t in targets],
                                           [t.gt_classes for t in targets], self.num_classes, self.proposal_matcher.thresholds[1],
                                           self.batch_size_per_image, self.positive_sample_fraction,
                                           seed_salt=getattr(self, "_drop_salt", None) if self.static_sampling else None,
                                           append_gt=append_gt, pad_background=self.static_sampling)
        if self.static_sampling:
            return self._static_samples(r, proposals), None, None
        counts = r["counts"].cpu().tolist()
        out = []
        for i, (p, t) in enumerate(zip(proposals, targets)):
            n = counts[i][1]
            idx = r["sampled_idx"][i, :n].long()
            q = p[idx]
            q.proposal_boxes = Boxes(r["boxes"][i, :n])
            q.gt_classes = r["classes"][i, :n]
            q.gt_boxes = Boxes(r["gt_boxes"][i, :n])
            out.append(q)
        n_fg = [c[0] for c in counts]
        n_bg = [c[1] - c[0] for c in counts]
        return out, n_fg, n_bg

    def _static_samples(self, r, proposals):
        """Fixed-shape form of the sampled batch: B rows per image, no host read and no per-image kernels — the sampled
        Instances are views of the kernel's padded outputs (they carry proposal_boxes, gt_classes, gt_boxes; the
        objectness logits, which nothing downstream of the sampler reads, are not gathered).  Padding rows (none, unless
        an image has fewer than B candidates — flagged) are labelled background by the kernel."""
        out = []
        for i, p in enumerate(proposals):
            q = Instances(p.image_size)
            q.proposal_boxes = Boxes(r["boxes"][i])
            q.gt_classes = r["classes"][i]
            q.gt_boxes = Boxes(r["gt_boxes"][i])
            out.append(q)
        self._defer("sample", r["counts"])
        return out

    def _defer(self, key, dev_tensor):
        """Asynchronous device -> pinned-host copy of a few logging scalars (a memcpy node when captured in a graph)."""
        buf = self._deferred.get(key)
        if buf is None or buf.numel() != dev_tensor.numel():
            buf = torch.zeros(dev_tensor.numel(), dtype=torch.float32).pin_memory()
            self._deferred[key] = buf
        buf.copy_(dev_tensor.detach().float().reshape(-1), non_blocking=True)

    def flush_deferred_logs(self):
        """Hand the scalars of the most recently completed step to the event storage (reference keys); raises if a
        static-sampling step did not have BATCH_SIZE_PER_IMAGE rows for every image."""
        st = get_event_storage()
        s = self._deferred.get("sample")
        if s is not None:
            c = s.reshape(-1, 2)
            n_fg, n_bg = float(c[:, 0].mean()), float((c[:, 1] - c[:, 0]).mean())
            bad = bool((c[:, 1] != self.batch_size_per_image).any())
            if bad:
                raise RuntimeError("STATIC_SAMPLING: an image had fewer than BATCH_SIZE_PER_IMAGE sampled proposals; "
                                   "run this batch with MODEL.B200.STATIC_SAMPLING = False")
            st.put_scalar("roi_head/num_fg_samples", n_fg)
            st.put_scalar("roi_head/num_bg_samples", n_bg)
        a = self._deferred.get("accuracy")
        if a is not None:
            acc, n_fg, fg_acc, fn, n = a.tolist()
            st.put_scalar("fast_rcnn/cls_accuracy", acc / max(n, 1))
            if n_fg > 0:
                st.put_scalar("fast_rcnn/fg_cls_accuracy", fg_acc / n_fg)
                st.put_scalar("fast_rcnn/false_negative", fn / n_fg)

    @torch.no_grad()
    def label_and_sample_proposals(self, proposals, targets):
        th, lb = self.proposal_matcher.thresholds, self.proposal_matcher.labels
        extra_gt = any(name.startswith("gt_") and name not in ("gt_boxes", "gt_classes") for t in targets for name in t.get_fields())
        on_device = (proposals and all(p.proposal_boxes.tensor.is_cuda for p in proposals) and len(th) == 3 and list(lb) == [0, 1] and
                     not extra_gt and max(len(p) + len(t) for p, t in zip(proposals, targets)) <= 4096 and
                     max(len(t) for t in targets) <= 256)
        in_kernel_append = self.proposal_append_gt and on_device and self.static_sampling
        if self.proposal_append_gt and not in_kernel_append:
            proposals = add_ground_truth_to_proposals([t.gt_boxes for t in targets], proposals)
        if on_device:
            out, n_fg, n_bg = self._label_and_sample_device(proposals, targets, append_gt=in_kernel_append)
            if n_fg is None:             # static sampling: the scalars are logged by flush_deferred_logs
                return out
            st = get_event_storage()
            st.put_scalar("roi_head/num_fg_samples", np.mean(n_fg))
            st.put_scalar("roi_head/num_bg_samples", np.mean(n_bg))
            return out
        out, n_fg, n_bg = [], [], []
        for p, t in zip(proposals, targets):
            has_gt = len(t) > 0
            idxs, labels = self.proposal_matcher(pairwise_iou(t.gt_boxes, p.proposal_boxes))
            sampled, gt_classes = self._sample_proposals(idxs, labels, t.gt_classes)
            p = p[sampled]
            p.gt_classes = gt_classes
            if has_gt:
                src = idxs[sampled]
                for name, value in t.get_fields().items():
                    if name.startswith("gt_") and not p.has(name):
                        p.set(name, value[src])
            else:
                p.gt_boxes = Boxes(t.gt_boxes.tensor.new_zeros((len(sampled), 4)))
            n_bg.append((gt_classes == self.num_classes).sum().item())
            n_fg.append(gt_classes.numel() - n_bg[-1])
            out.append(p)
        st = get_event_storage()
        st.put_scalar("roi_head/num_fg_samples", np.mean(n_fg))
        st.put_scalar("roi_head/num_bg_samples", np.mean(n_bg))
        return out

    def forward(self, images, features, proposals, targets=None):
        raise NotImplementedError()


@ROI_HEADS_REGISTRY.register()
class Res5ROIHeads(ROIHeads):
    """C4 head: shared ROIAlign + res5, then the predictor (plain DeFRCN baseline, no text)."""

    def __init__(self, cfg, input_shape):
        super().__init__(cfg, input_shape)
        assert len(self.in_features) == 1
        bh = cfg.MODEL.ROI_BOX_HEAD
        assert not cfg.MODEL.KEYPOINT_ON
        self.channels_last = bool(b200_opt(cfg, "CHANNELS_LAST", True))
        self.res5_dtype = getattr(torch, b200_opt(cfg, "RES5_DTYPE", "bfloat16"))
        self.skip_dead_bins = bool(b200_opt(cfg, "SKIP_DEAD_BINS", True))
        # frozen res5: "tcgen05" = every convolution on the CTA-pair GEMM kernel (res5_ops.py);
        # "cudnn" = the library path (layers._FrozenRes5MeanFn / forward_folded)
        self.res5_impl = os.environ.get("B200_RES5_IMPL") or str(b200_opt(cfg, "RES5_IMPL", "tcgen05"))
        self.pooler = ROIPooler(output_size=bh.POOLER_RESOLUTION, scales=(1.0 / self.feature_strides[self.in_features[0]],),
                                sampling_ratio=bh.POOLER_SAMPLING_RATIO, pooler_type=bh.POOLER_TYPE,
                                channels_last_out=self.channels_last)
        self.res5, self.out_channels = self._build_res5_block(cfg)
        self.output_layer = cfg.MODEL.ROI_HEADS.OUTPUT_LAYER
        self.box_predictor = ROI_HEADS_OUTPUT_REGISTRY.get(self.output_layer)(
            cfg, self.out_channels, self.num_classes, self.cls_agnostic_bbox_reg)

    def _build_res5_block(self, cfg):
        r = cfg.MODEL.RESNETS
        out_channels = r.RES2_OUT_CHANNELS * 8
        assert not r.DEFORM_ON_PER_STAGE[-1], "Deformable conv is not supported in the res5 head."
        blocks = make_stage(BottleneckBlock, 3, first_stride=2, in_channels=out_channels // 2,
                            bottleneck_channels=r.NUM_GROUPS * r.WIDTH_PER_GROUP * 8, out_channels=out_channels,
                            num_groups=r.NUM_GROUPS, norm=r.NORM, stride_in_1x1=r.STRIDE_IN_1X1)
        return nn.Sequential(*blocks), out_channels

    def _res5_forward(self, x, prestrided=False):
        """res5 stays on cuDNN.  Unless its weights are being trained, FrozenBN is folded into
        the convolutions and the stage runs in `res5_dtype` channels-last; gradients still flow to `x`."""
        if self._res5_trainable():
            return self.res5(x)
        y = x.to(self.res5_dtype)
        for i, blk in enumerate(self.res5):
            y = blk.forward_folded(y, prestrided=prestrided and i == 0)
        return y

    def _res5_trainable(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.res5.parameters())

    def _shared_roi_transform(self, features, boxes):
        # res5's first block reads the 7x7 pooled map through 1x1 stride-2 convolutions only (roi_heads.py:313-337 with
        # RESNETS.STRIDE_IN_1X1): 33 of 49 bins are dead.  On the frozen path the pooler emits just the live bins.
        blk0 = self.res5[0]
        skip = (self.skip_dead_bins and not self._res5_trainable() and blk0.reads_strided_1x1())
        step = blk0.stride if skip else 1
        return self._res5_forward(self.pooler(features, boxes, bin_step=step), prestrided=skip)

    def _res5_mean(self, x, prestrided=False):
        """res5 + mean over (h, w) of the pooled ROI map -> (R, C_out) fp32  [roi_heads.py:339-344 + :1109].
        Frozen res5 under autograd runs as one node (layers._FrozenRes5MeanFn: cuDNN convolutions, fused elementwise
        kernels); the spatial mean of a bf16 channels-last stage output is the C-ABI kernel in every mode."""
        from ... import res5_ops, train_ops
        frozen = not self._res5_trainable()
        if frozen and self.res5_impl == "tcgen05" and self.res5_dtype == torch.bfloat16 and x.is_cuda:
            pooled = res5_ops.frozen_res5_mean(self.res5, x.to(self.res5_dtype), prestrided)
            if pooled is not None:
                return pooled
        if frozen and torch.is_grad_enabled() and x.requires_grad and self.res5_dtype == torch.bfloat16:
            pooled = frozen_res5_mean(self.res5, x.to(self.res5_dtype), prestrided=prestrided)
            if pooled is not None:
                return pooled
        y = self._res5_forward(x, prestrided=prestrided)
        if (not y.requires_grad and y.is_cuda and y.dtype == torch.bfloat16 and y.shape[0] > 0 and y.shape[1] % 8 == 0 and
                y.permute(0, 2, 3, 1).is_contiguous()):
            return train_ops.spatial_mean(y)
        return y.mean(dim=[2, 3], dtype=torch.float32)

    def _pooled(self, features, proposals):
        # res5's first block reads the 7x7 pooled map through 1x1 stride-2 convolutions only (roi_heads.py:313-337 with
        # RESNETS.STRIDE_IN_1X1): 33 of 49 bins are dead.  On the frozen path the pooler emits just the live bins.
        blk0 = self.res5[0]
        skip = self.skip_dead_bins and not self._res5_trainable() and blk0.reads_strided_1x1()
        x = self.pooler([features[f] for f in self.in_features], [p.proposal_boxes for p in proposals],
                        bin_step=blk0.stride if skip else 1)
        self._mark("roi_align", x)
        pooled = self._res5_mean(x, prestrided=skip)
        self._after_res5_enqueued()
        self._mark("res5_mean", pooled)
        return pooled

    def _after_res5_enqueued(self):
        """Hook: res5 has been enqueued (not finished) — the place to start side-stream work that should share the GPU
        with res5's tensor-core kernels rather than with the pooler."""

    def forward(self, images, features, proposals, targets=None):
        del images
        if self.training:
            proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        feature_pooled = self._pooled(features, proposals)
        logits, deltas = self.box_predictor(feature_pooled)
        outputs = FastRCNNOutputs(self.box2box_transform, logits, deltas, proposals, self.smooth_l1_beta)
        if self.training:
            return [], outputs.losses()
        pred, _ = outputs.inference(self.test_score_thresh, self.test_nms_thresh, self.test_detections_per_img)
        return pred, {}


@ROI_HEADS_REGISTRY.register()
class SematicRes5ROIHeads(Res5ROIHeads):
    """C4 head whose classifier sees the ROI feature after cross-attention over class-name text embeddings."""

    def __init__(self, cfg, input_shape):
        super().__init__(cfg, input_shape)
        self.__init_LV_model__(self.out_channels, cfg)
        self.box_predictor = ROI_HEADS_OUTPUT_REGISTRY.get(self.output_layer)(
            cfg, self.out_channels, self.num_classes, self.cls_agnostic_bbox_reg)

    def __init_LV_model__(self, input_size, cfg):
        self.fused_training = bool(b200_opt(cfg, "FUSED_TRAINING", True))
        # optional cosine + temperature logits against the text prototypes (CrossOutput head); off = the reference's
        # un-normalised dot product (roi_heads.py:1157-1159)
        self.cosine_logits = bool(b200_opt(cfg, "COSINE_LOGITS", False))
        self.cosine_tau = float(b200_opt(cfg, "COSINE_TAU", 20.0))
        self.addition_model = cfg.MODEL.ADDITION.NAME
        self.semantic_dim = SEMANTIC_DIM[self.addition_model]
        self.attention = SematicProposalAttention(input_size, cfg=cfg, is_multi=False)
        if cfg.MODEL.ADDITION.FREEZEATTENTION:
            for p in self.attention.parameters():
                p.requires_grad = False
        self.output_projection = nn.Linear(input_size, self.semantic_dim)
        self.sematic_projection = nn.Linear(self.semantic_dim, input_size)
        self.projection_matrix = nn.Parameter(torch.randn(self.semantic_dim, input_size) * 1e-8)

    @torch.no_grad()
    def label_proposals(self, proposals, targets):
        """Attach GT labels to *all* proposals without sampling (test-with-GT mode).  Faithful to the reference,
        including its quirk: labels come out ordered [foreground..., background...] while the proposals keep
        their original order (roi_heads.py:1008-1028)."""
        out = []
        for p, t in zip(proposals, targets):
            idxs, labels = self.proposal_matcher(pairwise_iou(t.gt_boxes, p.proposal_boxes))
            gt_classes = self._assign_labels(idxs, labels, t.gt_classes)
            pos = nonzero_tuple((gt_classes != -1) & (gt_classes != self.num_classes))[0]
            neg = nonzero_tuple(gt_classes == self.num_classes)[0]
            order = torch.cat([pos, neg], dim=0)
            p.gt_classes = gt_classes[order]
            if len(t) > 0:
                src = idxs[order]
                for name, value in t.get_fields().items():
                    if name.startswith("gt_") and not p.has(name):
                        p.set(name, value[src])
            out.append(p)
        return out

    def cal_CE_att(self, output_att, gt_classes):
        a = F.relu(self.output_projection(output_att["sim2stext"]))
        score = F.softmax(torch.matmul(a, output_att["text_feat"].transpose(0, 1)), dim=1)
        return {"loss_attentive": F.cross_entropy(score, gt_classes, reduction="mean")}

    def forward_att(self, feature_pooled, gt_classes=0):
        attn, output_att = self.attention(feature_pooled)
        loss_att = {}
        if self.training:
            # CE over the attention *probabilities* (K+2 columns) — reference behaviour, roi_heads.py:1079-1081
            loss_att["loss_attentive"] = F.cross_entropy(attn[0], gt_classes, reduction="mean")
        logits, deltas = self.box_predictor(feature_pooled, output_att["sim2stext"],
                                            output_att.get("x_bf16"), output_att.get("sim2stext_bf16"))
        output_att["pred_logits"], output_att["pred_bbox"] = logits, deltas
        return output_att, loss_att

    _DROP_STEP = [0]

    def fused_train_losses(self, feature_pooled, proposals, gt_classes, teacher_logits=None, kd=None):
        """Fine-tune direction on the hand-written kernels (train_ops._FusedHeadTrain): text-fusion chain, predictor
        with classifier dropout, and the three losses in one autograd node; same numbers as `forward_att` +
        `FastRCNNOutputs.losses` (roi_heads.py:1060-1132) up to bf16 GEMM operands.  The (K+2)-row text side runs in
        fp32 on its own small kernels (train_ops._TextSide) and receives dKq / dVp from the fused node."""
        from ... import train_ops
        att, sa = self.attention, self.attention.attention
        pending, self._text_pending = getattr(self, "_text_pending", None), None
        if pending is None:
            kq, vp = train_ops.text_side(att)
        else:
            kq, vp, done = pending
            cur = torch.cuda.current_stream()
            cur.wait_event(done)
            kq.record_stream(cur)
            vp.record_stream(cur)
        # per-image views of the sampler's batched outputs sit back to back: no copy kernel then
        props = ops.cat_adjacent([p.proposal_boxes.tensor for p in proposals])
        gtb = ops.cat_adjacent([p.gt_boxes.tensor for p in proposals])
        pred = self.box_predictor
        drop = pred._dropout_ratio if pred._do_cls_dropout else 0.0
        salt = getattr(self, "_drop_salt", None)
        if salt is None:
            self._DROP_STEP[0] += 1
            seed = (torch.initial_seed() * 1000003 + self._DROP_STEP[0]) & 0x7FFFFFFFFFFFFFFF
        else:              # device-resident step counter (graph capture): the host part of the seed stays constant
            salt.add_(1)
            seed = (torch.initial_seed() * 1000003) & 0x7FFFFFFFFFFFFFFF
        cross = self._fused_cross_operands()
        losses, logits, self._acc_stats = train_ops.fused_head_train(
            feature_pooled, kq, vp, sa, pred, gt_classes, props, gtb, self.num_classes, self.box2box_transform.weights,
            self.smooth_l1_beta, drop, seed, cross is None, salt, teacher_logits, kd, cross)
        parts = train_ops.split_losses(losses)
        out = {"loss_cls": parts[0], "loss_box_reg": parts[1]}
        if cross is None:
            out["loss_attentive"] = parts[2]
        if teacher_logits is not None:
            out["loss_kl"] = parts[3]
        return out, logits

    def use_device_dropout_counter(self, enable=True):
        """Keep the classifier-dropout step counter in device memory (incremented by a kernel each step) instead of on the
        host, so that a CUDA graph captured over the training step draws a fresh mask on every replay."""
        dev = self.attention.attention.w_q.weight.device
        self._drop_salt = torch.zeros(1, dtype=torch.int64, device=dev) if enable else None

    def _fused_cross_operands(self):
        """None for the plain text-fused head; (output_projection, text prototypes) for the CrossOutput head."""
        return None

    def _fused_train_path(self):
        if not (self.training and torch.is_grad_enabled() and self.fused_training):
            return False
        if type(self).forward_att is SematicRes5ROIHeads.forward_att:
            return type(self.box_predictor).__name__ == "FastRCNNOutputLayers"
        # CrossOutput: prototype logits without the cosine option and without dropout on the logits
        return (type(self).forward_att is SematicRes5ROIHeadsCrossOutput.forward_att and not self.cosine_logits and
                type(self.box_predictor).__name__ == "FastRCNNAttentionOutputLayers" and
                not self.box_predictor._do_cls_dropout)

    def prefetch_text_side(self, after=None):
        """Start this step's text-side projections on a side stream (train_ops.text_side_async); `fused_train_losses`
        picks the result up.  Called once res5 has been enqueued, gated on an event from the top of `forward`, so that
        on the GPU they run under res5 (the slice-resident ROIAlign kernel is persistent with a static work split and
        needs whole SMs: sharing them costs its tail) while the host has nothing queued ahead of the pooler."""
        from ... import train_ops
        if self.attention.attention.w_q.weight.is_cuda:
            self._text_pending = train_ops.text_side_async(self.attention, after=after)

    def _after_res5_enqueued(self):
        if self._fused_train_path() and getattr(self, "_step_begin", None) is not None:
            self.prefetch_text_side(after=self._step_begin)
            self._step_begin = None

    def forward(self, images, features, proposals, targets=None):
        del images
        test_with_gt = (not self.training) and bool(targets)
        gt_classes = 0
        if self._fused_train_path() and self.attention.attention.w_q.weight.is_cuda:
            self._step_begin = torch.cuda.Event()          # the parameters hold this step's values from here on
            self._step_begin.record()
        if self.training:
            proposals = self.label_and_sample_proposals(proposals, targets)
            gt_classes = ops.cat_adjacent([p.gt_classes for p in proposals])
            self._mark("label_sample")
        elif test_with_gt:
            proposals = self.label_proposals(proposals, targets)
        feature_pooled = self._pooled(features, proposals)
        teacher_logits = self._teacher_logits(feature_pooled, gt_classes) if self.training else None
        if feature_pooled.is_cuda and self._fused_train_path():
            losses, logits = self.fused_train_losses(feature_pooled, proposals, gt_classes, teacher_logits,
                                                     self._kd_params() if teacher_logits is not None else None)
            # FastRCNNOutputs._log_accuracy (fast_rcnn.py:191-220): the counts come out of the loss kernel's pass
            if self.static_sampling:
                self._defer("accuracy", self._acc_stats)
            else:
                acc, n_fg, fg_acc, fn, n = self._acc_stats.tolist()
                st = get_event_storage()
                st.put_scalar("fast_rcnn/cls_accuracy", acc / max(n, 1))
                if n_fg > 0:
                    st.put_scalar("fast_rcnn/fg_cls_accuracy", fg_acc / n_fg)
                    st.put_scalar("fast_rcnn/false_negative", fn / n_fg)
            self._mark("text_fusion_losses")
            return [], losses
        att_output, att_loss = self.forward_att(feature_pooled, gt_classes)
        self._mark("text_fusion_predictor")
        outputs = FastRCNNOutputs(self.box2box_transform, att_output["pred_logits"], att_output["pred_bbox"], proposals,
                                  self.smooth_l1_beta)
        if self.training:
            losses = dict(outputs.losses())
            losses.update(att_loss)
            if teacher_logits is not None:
                from .my_module import loss_fn_kd_only
                T, alpha = self._kd_params()
                losses["loss_kl"] = loss_fn_kd_only(att_output["pred_logits"], gt_classes, self.num_classes, teacher_logits,
                                                    {"alpha": alpha, "temperature": T})
            return [], losses
        pred, _ = outputs.inference(self.test_score_thresh, self.test_nms_thresh, self.test_detections_per_img)
        self._mark("decode_nms")
        return pred, {}

    def _teacher_logits(self, feature_pooled, gt_classes):
        """Hook of the distillation head: logits of a frozen teacher for this batch, or None."""
        return None

    def _kd_params(self):
        return 1.0, 1.0


@ROI_HEADS_REGISTRY.register()
class SematicRes5ROIHeadsDistill(SematicRes5ROIHeads):
    """Student fine-tuning against a frozen, ground-truth-conditioned teacher (BASELINE configs[3]).

    The reference's distillation heads (`TextRes5ROIHeads*`, roi_heads.py:529-919) do not run as checked in
    (SURVEY §2.2); this is their intended flow (our_roi_heads_dnt.py:338-410, roi_heads.py:740-765): the teacher —
    `LV_attention_VKV` over the GT class embeddings plus its own classifier — sees the pooled ROI features and their
    labels, the student is the text-fused head, and `loss_kl = loss_fn_kd_only(student logits, gt, bg, teacher logits,
    alpha = 1, T = KL_TEMP)` (my_module.py:409-437; run_text_train_teacher_novel.sh: KL_TEMP 5) joins the student's own
    losses.  The teacher runs without autograd on the fused kernels (ops.teacher_attention_forward + the tcgen05 GEMM for
    its classifier); the student's step is the fused node with the KL term as a fourth loss."""

    def __init__(self, cfg, input_shape):
        super().__init__(cfg, input_shape)
        from .teacher_modules import LV_attention_VKV
        self.teacher = LV_attention_VKV(self.out_channels, cfg=cfg)
        self.teacher_cls_score = nn.Linear(self.out_channels, self.num_classes + 1)
        nn.init.normal_(self.teacher_cls_score.weight, std=0.01)
        nn.init.constant_(self.teacher_cls_score.bias, 0)
        for p in list(self.teacher.parameters()) + list(self.teacher_cls_score.parameters()):
            p.requires_grad = False
        self.kd_temp = float(b200_opt(cfg, "KD_TEMP", 5.0))
        self._teacher_w = None

    def _kd_params(self):
        return self.kd_temp, 1.0

    @torch.no_grad()
    def _teacher_logits(self, feature_pooled, gt_classes):
        _, out = self.teacher(feature_pooled.detach(), gt_classes)
        w, b = self.teacher_cls_score.weight, self.teacher_cls_score.bias
        if feature_pooled.is_cuda and "sim2stext_bf16" in out:
            key = (w.data_ptr(), w._version, ops.PARAM_GENERATION[0])
            if self._teacher_w is None or self._teacher_w[0] != key:
                self._teacher_w = (key, w.detach().to(torch.bfloat16).contiguous(), b.detach().float().contiguous())
            return ops.gemm_bf16(out["sim2stext_bf16"], self._teacher_w[1], self._teacher_w[2])
        return F.linear(out["sim2stext"][0], w, b)


@ROI_HEADS_REGISTRY.register()
class SematicRes5ROIHeadsCrossOutput(SematicRes5ROIHeads):
    """Logits are dot products of the projected fused feature with the text prototypes (roi_heads.py:1154-1171);
    used with OUTPUT_LAYER = FastRCNNAttentionOutputLayers."""

    def _fused_cross_operands(self):
        return self.output_projection, self.attention.forward_language_model()["text_feat"]

    def forward_att(self, feature_pooled, gt_classes=0):
        if self.training and torch.is_grad_enabled():
            _, output_att = self.attention(feature_pooled)
            a = F.relu(self.output_projection(output_att["sim2stext"]))
            t = output_att["text_feat"]
            if self.cosine_logits:         # my_module.py:461-469 `sim_matrix` x temperature (:449-458)
                a = a / a.norm(dim=1, keepdim=True).clamp_min(1e-12)
                t = t / t.norm(dim=1, keepdim=True).clamp_min(1e-12) * self.cosine_tau
            score = torch.matmul(a, t.transpose(0, 1))
            xb = None
        else:
            extra = {"output_projection.weight": self.output_projection.weight, "output_projection.bias": self.output_projection.bias}
            _, output_att = self.attention(feature_pooled, extra=extra)
            w = output_att["fused_w"]
            zb, wo, bo = output_att["sim2stext_bf16"], w["extra.output_projection.weight"], w["extra.output_projection.bias"]
            if self.cosine_logits:
                # my_module.py:449-469: x.t / (|x| |t|) x temperature.  |x|^2 is a by-product of the projection's epilogue, the
                # logits' epilogue divides by it; 1/|t| and the temperature sit in the (K+1)-row text operand — the
                # activations never take a normalisation pass
                ssq = torch.empty((zb.shape[0], -(-wo.shape[0] // 64)), dtype=torch.float32, device=zb.device)
                a = ops.gemm2(zb, wo, bias=bo, relu=True, rowsumsq_out=ssq)
                tb = ops.l2_normalize_rows(output_att["text_feat"].float().contiguous(), scale=self.cosine_tau)
                score = torch.empty((zb.shape[0], tb.shape[0]), dtype=torch.float32, device=zb.device)
                ops.gemm2(a, tb, row_scale_sumsq=ssq, out_f32=score, want_out=False)
            else:
                a = ops.gemm_bf16(zb, wo, bo, relu=True, out_dtype=torch.bfloat16)
                score = ops.gemm_bf16(a, output_att["text_feat"].to(torch.bfloat16).contiguous())
            xb = output_att.get("x_bf16")
        logits, deltas = self.box_predictor(feature_pooled, score, xb, None)
        output_att["pred_logits"], output_att["pred_bbox"] = logits, deltas
        return output_att, {}
