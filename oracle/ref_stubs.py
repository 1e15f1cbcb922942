"""oracle/ref_stubs.py — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Makes the *unmodified* reference sources under /root/reference importable in this container,
where the third-party packages they depend on are absent (detectron2==0.3, fvcore, torchnlp —
/root/reference/requirements.txt:13).  It does two things:

1. registers `sys.modules` stand-ins for `detectron2.*`, `fvcore.*`, `torchnlp.*` holding a
   minimal CPU restatement of the detectron2 v0.3 symbols the hot path touches (Boxes, Instances,
   Registry, Box2BoxTransform, ROIPooler -> torchvision.ops.roi_align(aligned=True),
   batched_nms -> the torchvision-0.8.1 coordinate-trick path, Matcher, subsample_labels,
   pairwise_iou, BottleneckBlock/make_stage with FrozenBN, get_event_storage, smooth_l1_loss);
2. registers empty *package shells* for `defrcn`, `defrcn.modeling`, ... with `__path__` pointing at
   /root/reference so that `import defrcn.modeling.roi_heads.fast_rcnn` executes the reference file
   itself but none of the `__init__.py` files that would drag in the engine/dataloader.

Used only by oracle/gen_golden.py (golden-vector generation, in-container) and by the tests that
validate the oracle port against the reference when /root/reference is present.  /root/reference
does not exist on the GPU box, so nothing GPU-marked depends on this file.
"""
import importlib
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

REFERENCE_ROOT = os.environ.get("B200ROI_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "defrcn"))


# --------------------------------------------------------------------------------------
# detectron2 v0.3 restatements (CPU, torch)
# --------------------------------------------------------------------------------------
class ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


def cat(tensors, dim=0):
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


class Registry:
    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._obj_map[o.__name__] = o
                return o
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return self._obj_map[name]


class Boxes:
    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((0, 4))
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device))

    def area(self):
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def clip(self, box_size):
        h, w = box_size
        self.tensor[:, 0].clamp_(min=0, max=w)
        self.tensor[:, 1].clamp_(min=0, max=h)
        self.tensor[:, 2].clamp_(min=0, max=w)
        self.tensor[:, 3].clamp_(min=0, max=h)

    def nonempty(self, threshold=0.0):
        # detectron2 0.3 structures/boxes.py::Boxes.nonempty
        box = self.tensor
        widths = box[:, 2] - box[:, 0]
        heights = box[:, 3] - box[:, 1]
        return (widths > threshold) & (heights > threshold)

    def scale(self, scale_x, scale_y):
        # detectron2 0.3 structures/boxes.py::Boxes.scale
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        return Boxes(b)

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device

    @classmethod
    def cat(cls, boxes_list):
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))


def pairwise_iou(boxes1, boxes2):
    area1, area2 = boxes1.area(), boxes2.area()
    b1, b2 = boxes1.tensor, boxes2.tensor
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    iou = torch.where(inter > 0, inter / (area1[:, None] + area2 - inter),
                      torch.zeros(1, dtype=inter.dtype, device=inter.device))
    return iou


class Instances:
    def __init__(self, image_size, **kwargs):
        self._image_size = image_size
        self._fields = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    @staticmethod
    def cat(instance_lists):
        ret = Instances(instance_lists[0].image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            ret.set(k, values)
        return ret


class Box2BoxTransform:
    def __init__(self, weights, scale_clamp=math.log(1000.0 / 16)):
        self.weights, self.scale_clamp = weights, scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        sw = src_boxes[:, 2] - src_boxes[:, 0]
        sh = src_boxes[:, 3] - src_boxes[:, 1]
        sx = src_boxes[:, 0] + 0.5 * sw
        sy = src_boxes[:, 1] + 0.5 * sh
        tw = target_boxes[:, 2] - target_boxes[:, 0]
        th = target_boxes[:, 3] - target_boxes[:, 1]
        tx = target_boxes[:, 0] + 0.5 * tw
        ty = target_boxes[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        dx = wx * (tx - sx) / sw
        dy = wy * (ty - sy) / sh
        dw = ww * torch.log(tw / sw)
        dh = wh * torch.log(th / sh)
        return torch.stack((dx, dy, dw, dh), dim=1)

    def apply_deltas(self, deltas, boxes):
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = deltas[:, 2::4] / ww
        dh = deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        pred_boxes = torch.zeros_like(deltas)
        pred_boxes[:, 0::4] = pred_ctr_x - 0.5 * pred_w
        pred_boxes[:, 1::4] = pred_ctr_y - 0.5 * pred_h
        pred_boxes[:, 2::4] = pred_ctr_x + 0.5 * pred_w
        pred_boxes[:, 3::4] = pred_ctr_y + 0.5 * pred_h
        return pred_boxes


def tv_nms(boxes, scores, iou_threshold):
    import torchvision
    return torchvision.ops.nms(boxes, scores, iou_threshold)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """detectron2 0.3 layers/nms.py::batched_nms over torchvision 0.8.1 boxes.py::batched_nms."""
    assert boxes.shape[-1] == 4
    if len(boxes) < 40000:
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device)
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
        return tv_nms(boxes_for_nms, scores, iou_threshold)
    result_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
    for cid in torch.unique(idxs).cpu().tolist():
        mask = (idxs == cid).nonzero().view(-1)
        keep = tv_nms(boxes[mask], scores[mask], iou_threshold)
        result_mask[mask[keep]] = True
    keep = result_mask.nonzero().view(-1)
    keep = keep[scores[keep].argsort(descending=True)]
    return keep


class ROIPooler(nn.Module):
    """Single-level detectron2 ROIPooler ('ROIAlignV2' == aligned=True, 'ROIAlign' == aligned=False)."""

    def __init__(self, output_size, scales, sampling_ratio, pooler_type,
                 canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(scales) == 1, "C4 heads use one level"
        assert pooler_type in ("ROIAlign", "ROIAlignV2")
        self.output_size, self.scale = output_size, scales[0]
        self.sampling_ratio, self.aligned = sampling_ratio, pooler_type == "ROIAlignV2"

    def forward(self, x, box_lists):
        import torchvision
        assert len(x) == 1
        rois = torch.cat([
            torch.cat([torch.full((len(b), 1), i, dtype=b.tensor.dtype, device=b.tensor.device),
                       b.tensor], dim=1) for i, b in enumerate(box_lists)], dim=0)
        return torchvision.ops.roi_align(x[0], rois.to(x[0].dtype), self.output_size, self.scale,
                                         self.sampling_ratio, self.aligned)


class Matcher:
    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        thresholds = thresholds[:]
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        self.thresholds, self.labels = thresholds, labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, mqm):
        if mqm.numel() == 0:
            default_matches = mqm.new_full((mqm.size(1),), 0, dtype=torch.int64)
            default_labels = mqm.new_full((mqm.size(1),), self.labels[0], dtype=torch.int8)
            return default_matches, default_labels
        matched_vals, matches = mqm.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for (l, low, high) in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            low_high = (matched_vals >= low) & (matched_vals < high)
            match_labels[low_high] = l
        return matches, match_labels


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = int(num_samples * positive_fraction)
    num_pos = min(positive.numel(), num_pos)
    num_neg = num_samples - num_pos
    num_neg = min(negative.numel(), num_neg)
    perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


def add_ground_truth_to_proposals(gt_boxes, proposals):
    out = []
    for gt, p in zip(gt_boxes, proposals):
        gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
        gt_logits = gt_logit_value * torch.ones(len(gt), device=gt.tensor.device)
        gp = Instances(p.image_size)
        gp.proposal_boxes = gt
        gp.objectness_logits = gt_logits
        out.append(Instances.cat([p, gp]))
    return out


class FrozenBatchNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        bias = self.bias - self.running_mean * scale
        return x * scale.reshape(1, -1, 1, 1).to(x.dtype) + bias.reshape(1, -1, 1, 1).to(x.dtype)


def get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    return {"FrozenBN": FrozenBatchNorm2d, "BN": nn.BatchNorm2d}[norm](out_channels)


class Conv2d(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm, self.activation = norm, activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, num_groups=1,
                 norm="BN", stride_in_1x1=False, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False,
                                   norm=get_norm(norm, out_channels))
        else:
            self.shortcut = None
        stride_1x1, stride_3x3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=stride_1x1,
                            bias=False, norm=get_norm(norm, bottleneck_channels))
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=stride_3x3,
                            padding=1 * dilation, bias=False, groups=num_groups, dilation=dilation,
                            norm=get_norm(norm, bottleneck_channels))
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False,
                            norm=get_norm(norm, out_channels))
        for layer in [self.conv1, self.conv2, self.conv3, self.shortcut]:
            if layer is not None:
                nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        shortcut = self.shortcut(x) if self.shortcut is not None else x
        out += shortcut
        return F.relu_(out)


def make_stage(block_class, num_blocks, first_stride, *, in_channels, out_channels, **kwargs):
    blocks = []
    for i in range(num_blocks):
        blocks.append(block_class(in_channels=in_channels, out_channels=out_channels,
                                  stride=first_stride if i == 0 else 1, **kwargs))
        in_channels = out_channels
    return blocks


class _EventStorage:
    def __init__(self):
        self.scalars = {}

    def put_scalar(self, name, value, smoothing_hint=True):
        self.scalars[name] = float(value)


_STORAGE = _EventStorage()


def get_event_storage():
    return _STORAGE


def smooth_l1_loss(input, target, beta, reduction="none"):
    """fvcore.nn.smooth_l1_loss."""
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        cond = n < beta
        loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


class ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor, self.image_sizes = tensor, image_sizes

    def __len__(self):
        return len(self.image_sizes)


class CfgNode(dict):
    """Attribute-style config node (enough of yacs for the keys the hot path reads)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def default_cfg(num_classes=20, addition="clip", train_dataset="voc_2007_trainval_all1_1shot_seed0",
                output_layer="FastRCNNOutputLayers", roi_head="SematicRes5ROIHeads"):
    """detectron2-0.3 defaults + defrcn/config/defaults.py:7-35 + main.py:36-44 for hot-path keys."""
    C = CfgNode
    cfg = C(
        MODEL=C(
            DEVICE="cpu", KEYPOINT_ON=False, MASK_ON=False,
            ROI_HEADS=C(NAME=roi_head, NUM_CLASSES=num_classes, BATCH_SIZE_PER_IMAGE=512,
                        POSITIVE_FRACTION=0.25, SCORE_THRESH_TEST=0.05, NMS_THRESH_TEST=0.5,
                        IN_FEATURES=["res4"], PROPOSAL_APPEND_GT=True, IOU_THRESHOLDS=[0.5],
                        IOU_LABELS=[0, 1], OUTPUT_LAYER=output_layer, CLS_DROPOUT=False,
                        DROPOUT_RATIO=0.8, ENABLE_DECOUPLE=True, BACKWARD_SCALE=0.001,
                        FREEZE_FEAT=False),
            ROI_BOX_HEAD=C(POOLER_RESOLUTION=7, POOLER_TYPE="ROIAlignV2", POOLER_SAMPLING_RATIO=0,
                           CLS_AGNOSTIC_BBOX_REG=False, SMOOTH_L1_BETA=0.0,
                           BBOX_REG_WEIGHTS=(10.0, 10.0, 5.0, 5.0), NAME=""),
            RESNETS=C(NUM_GROUPS=1, WIDTH_PER_GROUP=64, RES2_OUT_CHANNELS=256, STRIDE_IN_1X1=True,
                      NORM="FrozenBN", DEFORM_ON_PER_STAGE=[False, False, False, False], DEPTH=101),
            ADDITION=C(NAME=addition, INFERENCE_WITH_GT=False, TEACHER_TRAINING=False,
                       STUDENT_TRAINING=False, DISTIL_MODE=False, FREEZEATTENTION=False),
        ),
        TEST=C(DETECTIONS_PER_IMAGE=100, PCB_ENABLE=True, PCB_ALPHA=0.5, PCB_UPPER=1.0,
               PCB_LOWER=0.05, PCB_MODELTYPE="resnet", PCB_MODELPATH=""),
        DATASETS=C(TRAIN=(train_dataset,), TEST=("voc_2007_test_all1",)),
    )
    return cfg


# --------------------------------------------------------------------------------------
# sys.modules wiring
# --------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _pkg_shell(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


class FakeGloVe:
    """Stand-in for torchnlp GloVe (no vector files offline): a seeded 300-d vector per word."""

    def __init__(self, name="6B", dim=300):
        self.dim = dim

    def __getitem__(self, word):
        seed = sum((i + 1) * ord(c) for i, c in enumerate(word)) % (2 ** 31)
        return torch.randn(self.dim, generator=torch.Generator().manual_seed(seed))


VOC_ALL1 = ["aeroplane", "bicycle", "boat", "bottle", "car", "cat", "chair", "diningtable", "dog", "horse", "person",
            "pottedplant", "sheep", "train", "tvmonitor", "bird", "bus", "cow", "motorbike", "sofa"]


class _MetadataCatalog:
    def get(self, name):
        m = types.SimpleNamespace()
        m.thing_classes, m.base_classes, m.novel_classes = VOC_ALL1, VOC_ALL1[:15], VOC_ALL1[15:]
        m.novel_dataset_id_to_contiguous_id = {}
        return m


def glove_table(classes):
    """(K,300) class table the way LV_attention builds it (attentive_modules.py:352-364) from FakeGloVe."""
    map_voc = {"diningtable": "dining table", "pottedplant": "potted plant", "tvmonitor": "tv"}
    g = FakeGloVe()
    rows = []
    for name in classes:
        v = torch.zeros(300)
        for w in map_voc.get(name, name).split(" "):
            v = v + g[w]
        rows.append(v)
    return torch.stack(rows)


_installed = False


def install(class_embed_fn=None, device="cpu"):
    """Register stubs + package shells.  `class_embed_fn(class_names, model, include_bg)` replaces
    `get_class_embed` (the embedding .txt files are not in the reference repo)."""
    global _installed
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_ROOT)
    if _installed:
        return
    _mod("detectron2")
    _mod("detectron2.layers", ShapeSpec=ShapeSpec, cat=cat, nonzero_tuple=nonzero_tuple,
         batched_nms=batched_nms, Conv2d=Conv2d, get_norm=get_norm, FrozenBatchNorm2d=FrozenBatchNorm2d)
    _mod("detectron2.utils")
    _mod("detectron2.utils.registry", Registry=Registry)
    _mod("detectron2.utils.events", get_event_storage=get_event_storage)
    _mod("detectron2.structures", Boxes=Boxes, Instances=Instances, pairwise_iou=pairwise_iou,
         ImageList=ImageList)
    _mod("detectron2.modeling")
    _mod("detectron2.modeling.matcher", Matcher=Matcher)
    _mod("detectron2.modeling.poolers", ROIPooler=ROIPooler)
    _mod("detectron2.modeling.sampling", subsample_labels=subsample_labels)
    _mod("detectron2.modeling.box_regression", Box2BoxTransform=Box2BoxTransform)
    _mod("detectron2.modeling.backbone")
    _mod("detectron2.modeling.backbone.resnet", BottleneckBlock=BottleneckBlock, make_stage=make_stage)
    _mod("detectron2.modeling.proposal_generator")
    _mod("detectron2.modeling.proposal_generator.proposal_utils",
         add_ground_truth_to_proposals=add_ground_truth_to_proposals)
    _mod("detectron2.data", MetadataCatalog=_MetadataCatalog(), DatasetCatalog=object())
    _mod("fvcore")
    wi = _mod("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)
    _mod("fvcore.nn", smooth_l1_loss=smooth_l1_loss, weight_init=wi)
    _mod("torchnlp")
    _mod("torchnlp.word_to_vector", GloVe=FakeGloVe)

    r = os.path.join(REFERENCE_ROOT, "defrcn")
    _pkg_shell("defrcn", r)
    for sub in ("modeling", "modeling/roi_heads", "modeling/meta_arch", "modeling/proposal_generator", "utils", "data",
                "evaluation"):
        _pkg_shell("defrcn." + sub.replace("/", "."), os.path.join(r, sub))
    _installed = True

    # the reference hard-codes device='cuda' and reads datasets/{clip,glove}/*.txt: parametrise both
    ce = importlib.import_module("defrcn.utils.class_embedding")
    if class_embed_fn is not None:
        ce.get_class_embed = class_embed_fn


def load(modname):
    """Import a reference module by dotted name, e.g. 'defrcn.modeling.roi_heads.fast_rcnn'."""
    install()
    return importlib.import_module(modname)


def synthetic_class_embed(class_names, model, include_bg=False, seed=7):
    """unit-norm randn rows, seed 7 (the real CLIP/GloVe .txt files are not shipped)."""
    d = 512 if model == "clip" else 300
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(len(class_names) + (1 if include_bg else 0), d, generator=g)
    return (e / e.norm(dim=1, keepdim=True)).float()
