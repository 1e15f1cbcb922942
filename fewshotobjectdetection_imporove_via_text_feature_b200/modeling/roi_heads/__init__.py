from .fast_rcnn import (ROI_HEADS_OUTPUT_REGISTRY, FastRCNNAttentionOutputLayers, FastRCNNOutputLayers,
                        FastRCNNOutputs, fast_rcnn_inference, fast_rcnn_inference_single_image)
from .roi_heads import (ROI_HEADS_REGISTRY, Res5ROIHeads, ROIHeads, SematicRes5ROIHeads,
                        SematicRes5ROIHeadsCrossOutput, SematicRes5ROIHeadsDistill, build_roi_heads, select_foreground_proposals)
from .attentive_modules import (FFN, ScaledDotProductAttention, SematicProposalAttention, SingleHeadSiameseAttention)
from .my_module import loss_fn_kd, loss_fn_kd_only
from .teacher_modules import (LV_attention, LV_attention_textDomination, LV_attention_textDomination_VKV,
                              LV_attention_VKV)
