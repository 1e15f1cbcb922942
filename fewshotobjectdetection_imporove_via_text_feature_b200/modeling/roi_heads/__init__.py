from .fast_rcnn import (ROI_HEADS_OUTPUT_REGISTRY, FastRCNNAttentionOutputLayers, FastRCNNOutputLayers,
                        FastRCNNOutputs, fast_rcnn_inference, fast_rcnn_inference_single_image)
from .roi_heads import (ROI_HEADS_REGISTRY, Res5ROIHeads, ROIHeads, SematicRes5ROIHeads,
                        SematicRes5ROIHeadsCrossOutput, build_roi_heads, select_foreground_proposals)
