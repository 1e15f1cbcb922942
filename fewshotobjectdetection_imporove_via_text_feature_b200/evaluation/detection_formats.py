"""Detection records in the evaluators' formats, straight from the padded device tensors.

Reference: `PascalVOCDetectionEvaluator.process` (defrcn/evaluation/pascal_voc_evaluation.py:57-78) builds, per
detection, the VOC text line "{image_id} {score:.3f} {xmin:.1f} {ymin:.1f} {xmax:.1f} {ymax:.1f}" (xmin, ymin + 1 in
fp32) and, through `instances_to_coco_json` (:172-203, coco_evaluation.py:244-274), the COCO dict with an XYWH box
(fp32 subtraction).  The reference moves every image's `Instances` to the CPU field by field; here the whole padded
batch comes back in ONE device->host copy (`pack_batch`), and the strings / dicts are produced from it with the
reference's arithmetic, digit for digit.
"""
from collections import defaultdict

import numpy as np
import torch


def pack_batch(det):
    """dict(boxes (N,T,4), scores (N,T), classes (N,T) int64, counts (N)) on the device -> host tuple
    (boxes fp32 (N,T,4), scores fp32 (N,T), classes int64 (N,T), counts int (N)), one transfer."""
    N, T = det["scores"].shape
    buf = torch.empty((N, T * 6 + 1), dtype=torch.float32, device=det["scores"].device)
    buf[:, :T * 4] = det["boxes"].reshape(N, T * 4)
    buf[:, T * 4:T * 5] = det["scores"]
    buf[:, T * 5:T * 6] = det["classes"].to(torch.float32)          # class ids are far below 2^24: exact
    buf[:, T * 6] = det["counts"].to(torch.float32)
    h = buf.cpu().numpy()
    return (h[:, :T * 4].reshape(N, T, 4).copy(), h[:, T * 4:T * 5].copy(), h[:, T * 5:T * 6].astype(np.int64),
            h[:, T * 6].astype(np.int64))


def voc_prediction_lines(image_ids, boxes, scores, classes, counts, predictions=None):
    """-> {class id: [line, ...]} exactly as pascal_voc_evaluation.py:57-70 accumulates `self._predictions`."""
    predictions = defaultdict(list) if predictions is None else predictions
    for image_id, b, s, c, n in zip(image_ids, boxes, scores, classes, counts):
        b = np.array(b[:n], dtype=np.float32)
        b[:, 0] += 1                                                  # fp32 adds, like `xmin += 1` on np.float32 scalars
        b[:, 1] += 1
        for (xmin, ymin, xmax, ymax), score, cls in zip(b, s[:n].tolist(), c[:n].tolist()):
            predictions[cls].append(f"{image_id} {score:.3f} {xmin:.1f} {ymin:.1f} {xmax:.1f} {ymax:.1f}")
    return predictions


def coco_json_records(image_ids, boxes, scores, classes, counts):
    """-> per image list of {"image_id", "category_id", "bbox" [x, y, w, h], "score"} (instances_to_coco_json)."""
    out = []
    for image_id, b, s, c, n in zip(image_ids, boxes, scores, classes, counts):
        b = np.array(b[:n], dtype=np.float32)
        b[:, 2] -= b[:, 0]                                            # BoxMode.convert XYXY_ABS -> XYWH_ABS, fp32
        b[:, 3] -= b[:, 1]
        out.append([{"image_id": image_id, "category_id": cls, "bbox": box, "score": score}
                    for box, score, cls in zip(b.tolist(), s[:n].tolist(), c[:n].tolist())])
    return out
