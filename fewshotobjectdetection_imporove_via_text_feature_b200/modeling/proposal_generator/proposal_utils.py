"""RPN proposal selection — the stage that feeds the ROI head.

Mirror of `find_top_rpn_proposals` (reference: defrcn/modeling/proposal_generator/proposal_utils.py:13-118, the vendored
detectron2 0.3 function `GeneralizedRCNN`'s proposal generator ends with): same signature, same `list[Instances]`
result (`proposal_boxes`, `objectness_logits`, sorted by objectness).  The reference sorts every anchor's logit,
indexes, filters and runs `batched_nms` image by image with a host synchronisation per image
(`keep.sum().item()`, torchvision's mask read-back); here the whole batch is one C-ABI call
(`b200_rpn_select_proposals`: radix select + shared-memory sort, ordered filter, presorted per-level NMS) and the host
reads the per-image counts once at the end to slice the padded result.
"""
from typing import List, Tuple

import torch

from ... import ops
from ...structures import Boxes, Instances


def find_top_rpn_proposals_device(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                                  image_sizes: List[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int,
                                  post_nms_topk: int, min_box_size: float):
    """No-synchronisation form: padded `dict(boxes (N,post,4), logits (N,post), counts (N), n_invalid (N))` on the
    device — what a fused trainer feeds straight into `label_and_sample_proposals` / the pooler."""
    level_sizes = [int(l.shape[1]) for l in pred_objectness_logits]
    if len(proposals) == 1:
        boxes, logits = proposals[0], pred_objectness_logits[0]
    else:
        boxes, logits = torch.cat(list(proposals), dim=1), torch.cat(list(pred_objectness_logits), dim=1)
    image_hw = ops.image_hw_tensor(image_sizes, logits.device)
    return ops.rpn_select_proposals(boxes, logits, level_sizes, image_hw, nms_thresh, pre_nms_topk, post_nms_topk,
                                    min_box_size)


def find_top_rpn_proposals(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                           image_sizes: List[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int, post_nms_topk: int,
                           min_box_size: float, training: bool):
    """Drop-in for proposal_utils.py:13-118.  `proposals[l]` (N, Hl*Wl*A, 4), `pred_objectness_logits[l]` (N, Hl*Wl*A)."""
    out = find_top_rpn_proposals_device(proposals, pred_objectness_logits, image_sizes, nms_thresh, pre_nms_topk,
                                        post_nms_topk, min_box_size)
    host = torch.stack([out["counts"], out["n_invalid"]]).tolist()      # the one device->host read of the call
    if training and any(v != 0 for v in host[1]):
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    results = []
    for n, image_size in enumerate(image_sizes):
        res = Instances(image_size)
        res.proposal_boxes = Boxes(out["boxes"][n, :host[0][n]])
        res.objectness_logits = out["logits"][n, :host[0][n]]
        results.append(res)
    return results
