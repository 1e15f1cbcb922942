"""Detection post-processing to the requested output resolution.

Mirror of detectron2 0.3 `detector_postprocess`, which the reference calls right after the ROI head
(defrcn/modeling/meta_arch/rcnn.py:69-73): rescale the boxes from the network input size to
`(output_height, output_width)`, clip, drop empty boxes.  `detector_postprocess_batch` does it for the whole padded
batch of `fast_rcnn_inference_device` in one launch without leaving the device.
"""
import torch

from .. import ops
from ..structures import Boxes, Instances


def detector_postprocess_batch(det, image_sizes, output_sizes):
    """`det`: dict from `ops.fast_rcnn_inference_device` (boxes, scores, classes, roi_inds, counts); updated in place."""
    counts = det["counts"].contiguous()
    ops.detector_postprocess_(det["boxes"], det["scores"], det["classes"], det.get("roi_inds"), counts, image_sizes,
                              output_sizes)
    det["counts"] = counts
    return det


def detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    """Drop-in for detectron2.modeling.postprocessing.detector_postprocess on box-only `Instances` (the C4 head has no
    masks / keypoints)."""
    if results.has("pred_boxes"):
        name = "pred_boxes"
    elif results.has("proposal_boxes"):
        name = "proposal_boxes"
    else:
        raise KeyError("detector_postprocess: Instances hold neither pred_boxes nor proposal_boxes")
    n = len(results)
    fields = dict(results.get_fields())
    boxes = fields.pop(name).tensor.float().contiguous().clone().view(1, n, 4)
    dev = boxes.device
    order = torch.arange(n, dtype=torch.int64, device=dev).view(1, n)          # rides along: the surviving rows
    scores = torch.zeros((1, n), dtype=torch.float32, device=dev)
    counts = torch.full((1,), n, dtype=torch.int32, device=dev)
    if n:
        ops.detector_postprocess_(boxes, scores, order, None, counts, [results.image_size],
                                  [(output_height, output_width)])
    k = int(counts.item())
    keep = order[0, :k]
    out = Instances((output_height, output_width))
    out.set(name, Boxes(boxes[0, :k]))
    for f, v in fields.items():
        out.set(f, v[keep])
    return out
