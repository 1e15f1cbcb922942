"""Prototypical Calibration Block (PCB), B200-native.

Mirror of defrcn/evaluation/calibration_layer.py:17-139.  The reference re-reads the image with cv2, runs an
ImageNet ResNet-101, pools 1x1 ROI features, and then loops over detections in Python calling sklearn's
cosine_similarity on the CPU (one `.cpu().numpy()` per detection).  Here the pooling is the ROIAlign kernel
(1x1 @ 1/32, adaptive sampling) and the cosine + blend is one kernel over all detections of the image
(csrc/fusion_elem.cu::pcb_cosine_blend_kernel); scores are updated in place and — like the reference — are not
re-sorted.

`PrototypicalCalibrationBlock(cfg)` — the reference's one-argument call (evaluator.py:90) — builds the ImageNet ResNet-101
(evaluation/archs/resnet.py; weights from cfg.TEST.PCB_MODELPATH), the reference's preprocessing (calibration_layer.py:91-98:
BGR uint8 -> /255 -> (x - mean) / std -> RGB) and cv2 as the image reader.  What the reference additionally pulls from its
dataloader — the support set the prototypes are averaged from (`build_detection_test_loader(cfg, cfg.DATASETS.TRAIN[0])`,
outside the hot path) — is handed over instead: `support=` (an iterable of the reference's dataset dicts, or of
(image, Boxes, gt_classes) triples), `build_prototypes(support)` later, or ready-made `prototypes=`.  Any callable
`feature_extractor(image_bgr_uint8_hwc) -> (1,C,H/32,W/32)` plus an `fc` module can replace the CNN.
"""
import logging
import os

import torch

from .. import ops
from ..modeling.poolers import ROIPooler
from ..structures import Boxes

COCO_BASE_EXCLUDE = [7, 9, 10, 11, 12, 13, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38,
                     40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 59, 61, 63, 64, 65, 66, 67, 68, 69,
                     70, 71, 72, 73, 74, 75, 76, 77, 78, 79]


logger = logging.getLogger(__name__)

_MEAN_BGR, _STD_BGR = (0.406, 0.456, 0.485), (0.225, 0.224, 0.229)      # calibration_layer.py:91-92


def _read_bgr(path):
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return img


class PrototypicalCalibrationBlock:
    def __init__(self, cfg, feature_extractor=None, fc=None, prototypes=None, image_reader=None, support=None):
        self.cfg = cfg
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.alpha = cfg.TEST.PCB_ALPHA
        self.imagenet_model = None
        if feature_extractor is None:            # the reference's own construction (calibration_layer.py:24,31-42)
            self.imagenet_model = self.build_model()
            feature_extractor = self._imagenet_features
            fc = self.imagenet_model.fc if fc is None else fc
        self.feature_extractor, self.fc = feature_extractor, fc
        self.image_reader = _read_bgr if image_reader is None else image_reader
        self.roi_pooler = ROIPooler(output_size=(1, 1), scales=(1 / 32,), sampling_ratio=0, pooler_type="ROIAlignV2")
        self.exclude_cls = self.clsid_filter()
        self.prototypes = {}
        self._proto_mat = self._exclude_mask = None
        if prototypes is not None:
            self.set_prototypes(prototypes)
        elif support is not None:
            self.build_prototypes(support)

    def build_model(self):
        """ImageNet ResNet-101 in eval mode with the weights of cfg.TEST.PCB_MODELPATH (calibration_layer.py:31-42)."""
        from .archs import resnet101
        if self.cfg.TEST.PCB_MODELTYPE != "resnet":
            raise NotImplementedError(self.cfg.TEST.PCB_MODELTYPE)
        model = resnet101()
        path = self.cfg.TEST.PCB_MODELPATH
        if path and os.path.exists(path):
            logger.info("Loading ImageNet Pre-train Model from %s", path)
            model.load_state_dict(torch.load(path, map_location="cpu"))
        else:
            logger.warning("PCB: cfg.TEST.PCB_MODELPATH (%r) not found — the ImageNet extractor keeps its random initialisation", path)
        return model.to(self.device).eval()

    @torch.no_grad()
    def _imagenet_features(self, img):
        """calibration_layer.py:91-98: BGR uint8 (H,W,3) -> normalised RGB batch of one -> layer4 feature (1,2048,H/32,W/32)."""
        x = torch.as_tensor(img).to(self.device).permute(2, 0, 1).float()
        mean = torch.tensor(_MEAN_BGR, device=self.device).reshape(3, 1, 1)
        std = torch.tensor(_STD_BGR, device=self.device).reshape(3, 1, 1)
        x = ((x / 255.0 - mean) / std)[[2, 1, 0]].unsqueeze(0)
        return self.imagenet_model(x)[1]

    # -- prototypes -----------------------------------------------------------------------------------
    def set_prototypes(self, prototypes):
        """prototypes: {class_id: (1,D) tensor} (the reference's dict) or a dense (K,D) tensor."""
        if isinstance(prototypes, dict):
            self.prototypes = prototypes
            K = max(prototypes) + 1
            D = next(iter(prototypes.values())).shape[-1]
            mat = torch.zeros(K, D)
            missing = torch.ones(K, dtype=torch.uint8)
            for c, p in prototypes.items():
                mat[c] = p.reshape(-1).float().cpu()
                missing[c] = 0
        else:
            mat = prototypes.float().cpu()
            self.prototypes = {c: mat[c:c + 1] for c in range(mat.shape[0])}
            missing = torch.zeros(mat.shape[0], dtype=torch.uint8)
        for c in self.exclude_cls:
            if c < missing.numel():
                missing[c] = 1
        self._proto_mat = mat.to(self.device)
        self._exclude_mask = missing.to(self.device)

    def build_prototypes(self, support):
        """support: iterable of (image, Boxes, gt_classes).  Class prototype = mean ROI feature (:44-82)."""
        feats, labels = [], []
        for item in support:
            if isinstance(item, dict):      # the reference's dataset dict (calibration_layer.py:48-57): boxes are in the
                img = self.image_reader(item["file_name"])       # resized frame, the image is read at its file size
                inst = item["instances"]
                ratio = img.shape[0] / inst.image_size[0]
                boxes, gt = Boxes(inst.gt_boxes.tensor * ratio), inst.gt_classes
            else:
                img, boxes, gt = item
            feats.append(self.extract_roi_features(img, [boxes]).float().cpu())
            labels.append(gt.cpu())
        feats, labels = torch.cat(feats), torch.cat(labels)
        protos = {int(c): feats[labels == c].mean(dim=0, keepdim=True) for c in labels.unique().tolist()}
        self.set_prototypes(protos)
        return protos

    # -- features -------------------------------------------------------------------------------------
    @torch.no_grad()
    def extract_roi_features(self, img, boxes):
        """img: HxWx3 BGR uint8 (numpy or tensor) or a precomputed (1,C,h,w) conv feature.  The extractor is frozen: no graph."""
        if torch.is_tensor(img) and img.dim() == 4:
            conv_feature = img.to(self.device)
        else:
            conv_feature = self.feature_extractor(img)
        boxes = [b if isinstance(b, Boxes) or hasattr(b, "tensor") else Boxes(b) for b in boxes]
        boxes = [Boxes(b.tensor.to(self.device)) for b in boxes]
        pooled = self.roi_pooler([conv_feature.float()], boxes).flatten(1)
        return self.fc(pooled) if self.fc is not None else pooled

    # -- calibration ----------------------------------------------------------------------------------
    def execute_calibration(self, inputs, dts):
        inst = dts[0]["instances"]
        n = len(inst)
        if n == 0:
            return dts
        img = inputs[0].get("conv_feature") if isinstance(inputs[0], dict) else None
        if img is None:
            img = self.image_reader(inputs[0]["file_name"])
        # features for every detection; the kernel applies the [ileft, iright) score window itself, on the device
        feats = self.extract_roi_features(img, [inst.pred_boxes])
        scores = inst.scores if inst.scores.is_contiguous() else inst.scores.contiguous()
        ops.pcb_cosine_blend_(scores, feats, self._proto_mat, inst.pred_classes, self._exclude_mask, self.alpha,
                              self.cfg.TEST.PCB_LOWER, self.cfg.TEST.PCB_UPPER)
        if scores is not inst.scores:
            inst.scores.copy_(scores)
        return dts

    def clsid_filter(self):
        dsname = self.cfg.DATASETS.TEST[0]
        if "test_all" in dsname:
            if "coco" in dsname:
                return list(COCO_BASE_EXCLUDE)
            if "voc" in dsname:
                return list(range(0, 15))
            raise NotImplementedError
        return []
