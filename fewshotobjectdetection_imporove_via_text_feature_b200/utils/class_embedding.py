"""Class-name text embeddings (CLIP 512-d / GloVe 300-d) for the text-fused head.

Input contract of defrcn/utils/class_embedding.py:4-24 and class_name.py:4-23: one `.txt` vector per class
under `<root>/{clip,glove}/<class>.txt`.  The embedding files are not part of the reference repository, so a
seeded synthetic generator (unit-norm randn rows, seed 7) stands in when they are absent.
Unlike the reference, nothing here hard-codes 'cuda': tensors are created on CPU and moved by the module.
"""
import os
import re

import numpy as np
import torch

SEMANTIC_DIM = {"clip": 512, "glove": 300}

PASCAL_VOC_ALL_CATEGORIES = {
    1: ["aeroplane", "bicycle", "boat", "bottle", "car", "cat", "chair", "diningtable", "dog", "horse", "person",
        "pottedplant", "sheep", "train", "tvmonitor", "bird", "bus", "cow", "motorbike", "sofa"],
    2: ["bicycle", "bird", "boat", "bus", "car", "cat", "chair", "diningtable", "dog", "motorbike", "person",
        "pottedplant", "sheep", "train", "tvmonitor", "aeroplane", "bottle", "cow", "horse", "sofa"],
    3: ["aeroplane", "bicycle", "bird", "bottle", "bus", "car", "chair", "cow", "diningtable", "dog", "horse",
        "person", "pottedplant", "train", "tvmonitor", "boat", "cat", "motorbike", "sheep", "sofa"],
}
PASCAL_VOC_NOVEL_CATEGORIES = {k: v[15:] for k, v in PASCAL_VOC_ALL_CATEGORIES.items()}
PASCAL_VOC_BASE_CATEGORIES = {k: v[:15] for k, v in PASCAL_VOC_ALL_CATEGORIES.items()}


def get_class_name(cfg):
    """Class list from `cfg.DATASETS.TRAIN[0]` (voc_*_{base,novel,all}<split>...); COCO and unknown names fall
    back to generic `class_<i>` labels of length NUM_CLASSES (the COCO table lives in the data layer, out of scope)."""
    name = cfg.DATASETS.TRAIN[0]
    k = cfg.MODEL.ROI_HEADS.NUM_CLASSES
    classes = None
    if "voc" in name:
        # the reference indexes fixed token positions (class_name.py:8-13), which only works for names shaped like
        # `..._base1` / `..._all1_1shot_seed0`; the split id is taken from the `<kind><id>` token wherever it is
        m = re.search(r"(base|novel|all)(\d)", name)
        if m:
            table = {"base": PASCAL_VOC_BASE_CATEGORIES, "novel": PASCAL_VOC_NOVEL_CATEGORIES, "all": PASCAL_VOC_ALL_CATEGORIES}
            classes = table[m.group(1)].get(int(m.group(2)))
    if classes is None or len(classes) != k:
        classes = ["class_%d" % i for i in range(k)]
    return list(classes)


def synthetic_class_embed(class_names, model, include_bg=False, seed=7):
    d = SEMANTIC_DIM[model]
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(len(class_names) + (1 if include_bg else 0), d, generator=g)
    return (e / e.norm(dim=1, keepdim=True)).float()


def get_class_embed(class_names, model, include_bg=False, root="datasets"):
    files = [os.path.join(root, model, "%s.txt" % c) for c in class_names]
    if include_bg:
        files.append(os.path.join(root, model, "background.txt"))
    if all(os.path.exists(f) for f in files):
        return torch.tensor(np.array([np.loadtxt(f) for f in files])).float()
    return synthetic_class_embed(class_names, model, include_bg)


def create_normalized_orthogonal_tensor(tensor, generator=None):
    """bg = normalise(t - <t, r> r), r ~ N(0,1)  (reference: class_embedding.py:15-24)."""
    r = torch.randn(tensor.shape, generator=generator, dtype=tensor.dtype)
    r = r.to(tensor.device)
    o = tensor - torch.dot(tensor.flatten(), r.flatten()) * r
    return o / torch.norm(o)
