"""Gradient Decoupled Layer + AffineLayer, B200-native.

Mirrors defrcn/modeling/meta_arch/gdl.py:6-38 (same names, same parameter shapes `(1,C,1,1)`, so
`affine_rcnn.{weight,bias}` load unchanged).  On CUDA both directions run as ONE fused kernel each
(csrc/gdl_affine.cu) instead of the reference's identity-autograd-function + mul + add kernels, and the
output can be produced directly in the layout/dtype the ROIAlign gather wants.
"""
import torch
from torch import nn

from ... import ops


class GradientDecoupleLayer(torch.autograd.Function):
    """Identity forward, grad * lambda backward (CUDA 4-d maps: csrc/gdl_affine.cu with weight=NULL)."""

    @staticmethod
    def forward(ctx, x, _lambda):
        ctx._lambda = _lambda
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        if grad_output.is_cuda and grad_output.dim() == 4 and grad_output.dtype in (torch.float32, torch.bfloat16):
            return ops.gdl_scale(grad_output, ctx._lambda), None
        return grad_output * ctx._lambda, None


def decouple_layer(x, _lambda):
    return GradientDecoupleLayer.apply(x, _lambda)


class AffineLayer(nn.Module):
    def __init__(self, num_channels, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(1, num_channels, 1, 1))
        self.bias = nn.Parameter(torch.zeros(1, num_channels, 1, 1)) if bias else None

    def forward(self, X, _lambda=None, channels_last_out=False, out_dtype=None):
        """`affine(X)`; with `_lambda` given it is `affine(decouple_layer(X, _lambda))` in one fused pass
        (rcnn.py:94-97 calls the two back to back)."""
        if X.is_cuda:
            return ops.gdl_affine(X, self.weight, self.bias, 1.0 if _lambda is None else _lambda, out_dtype,
                                  channels_last_out)
        raise RuntimeError("b200roi AffineLayer runs on CUDA only (no CPU fallback)")


def decoupled_affine(x, affine, _lambda, channels_last_out=False, out_dtype=None):
    """Fused replacement for `affine(decouple_layer(x, _lambda))`."""
    return affine(x, _lambda, channels_last_out, out_dtype)
