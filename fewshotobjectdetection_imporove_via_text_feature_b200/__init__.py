"""B200-native (sm_100a) ROI-head hot path of the DeFRCN text-fused few-shot detector.

Public surface (mirrors the reference's `defrcn.modeling` / `defrcn.evaluation` names for this path):
    modeling.build_roi_heads, modeling.ROI_HEADS_REGISTRY, modeling.ROI_HEADS_OUTPUT_REGISTRY,
    modeling.{Res5ROIHeads, SematicRes5ROIHeads, SematicRes5ROIHeadsCrossOutput},
    modeling.{FastRCNNOutputs, FastRCNNOutputLayers, FastRCNNAttentionOutputLayers},
    modeling.{AffineLayer, decouple_layer, decoupled_affine},
    evaluation.PrototypicalCalibrationBlock, ops.*, distributed.*
Importing the package does not load the CUDA library; the first op does, and raises if it is missing.
"""
from . import config, structures  # noqa: F401

__all__ = ["config", "structures", "ops", "modeling", "evaluation", "distributed"]
__version__ = "0.1.0"
