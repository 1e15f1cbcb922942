from .meta_arch.gdl import AffineLayer, GradientDecoupleLayer, decouple_layer, decoupled_affine
from .roi_heads import (
    ROI_HEADS_OUTPUT_REGISTRY,
    ROI_HEADS_REGISTRY,
    FastRCNNAttentionOutputLayers,
    FastRCNNOutputLayers,
    FastRCNNOutputs,
    Res5ROIHeads,
    ROIHeads,
    SematicRes5ROIHeads,
    SematicRes5ROIHeadsCrossOutput,
    build_roi_heads,
)
