"""Single-level ROIPooler (detectron2.modeling.poolers.ROIPooler as configured at
defrcn/modeling/roi_heads/roi_heads.py:300-305 and calibration_layer.py:27): 'ROIAlignV2' == aligned=True,
'ROIAlign' == aligned=False, `sampling_ratio=0` == adaptive grid.  Runs csrc/roi_align.cu."""
from torch import nn

from .. import ops
from ..structures import Boxes


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4,
                 channels_last_out=False):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        if len(scales) != 1:
            raise NotImplementedError("the C4 ROI head pools from a single level (res4)")
        if pooler_type not in ("ROIAlign", "ROIAlignV2"):
            raise ValueError("unsupported pooler type %s" % pooler_type)
        self.output_size, self.scale = tuple(output_size), float(scales[0])
        self.sampling_ratio, self.aligned = int(sampling_ratio), pooler_type == "ROIAlignV2"
        self.channels_last_out = channels_last_out

    def forward(self, x, box_lists, bin_step=1):
        """bin_step=s returns only the bins [::s, ::s] of the pooled map (dead-bin skipping for a stride-s consumer)."""
        assert isinstance(x, list) and len(x) == 1, "ROIPooler expects a single-level feature list"
        tensors = [b.tensor if isinstance(b, Boxes) or hasattr(b, "tensor") else b for b in box_lists]
        rois, offsets = ops.boxes_to_rois(tensors)
        return ops.roi_align(x[0], rois, self.output_size, self.scale, self.sampling_ratio, self.aligned,
                             self.channels_last_out, offsets, bin_step)
