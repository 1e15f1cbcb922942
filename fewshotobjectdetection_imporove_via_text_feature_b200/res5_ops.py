"""Frozen res5 stage + spatial mean on the CTA-pair tcgen05 GEMM (csrc/gemm2_tcgen05.cu) — SURVEY §8f-1.

Reference: `Res5ROIHeads._build_res5_block` / `_shared_roi_transform` (defrcn/modeling/roi_heads/roi_heads.py:313-344:
three detectron2 BottleneckBlocks 1024 -> 2048, stride 2 in the first 1x1, FrozenBN) followed by
`box_features.mean(dim=[2, 3])` (:1109), under ROI_HEADS.FREEZE_FEAT (the res5 weights take no gradient).

With the weights frozen, FrozenBN folds into the convolutions and every convolution of the stage is a GEMM over the
NHWC pixel rows of the pooled ROI map (R ROIs x 4x4 pixels = 16 R rows):
    conv1 / conv3 / shortcut (1x1)   plain GEMM, K = C_in
    conv2 (3x3, pad 1)               implicit GEMM, K = 9 x 512, A fetched tap by tap by TMA (zero fill = padding)
    conv3 + shortcut (block 0)       ONE GEMM over K = 512 + 1024 (two A tensors, one accumulator)
    conv3 + identity (blocks 1, 2)   residual added in the epilogue
    bias (folded BN shift) + ReLU    epilogue; the ReLU masks the backward needs leave as packed bits (1 bit / element)
    mean over the 4x4 pixels         epilogue of the last conv3 (the last activation itself is never written)
Backward (data gradient only): the same kernel with transposed / tap-flipped frozen weights; ReLU backward reads the
packed masks in the epilogue, the residual fan-in of the two branches is the epilogue's residual operand (blocks 1, 2)
or a second K segment (block 0: conv1 and shortcut gradients in one accumulator).  No elementwise passes remain
between the GEMMs except the broadcast of the pooled gradient (`b200_mean_bwd_relu_bits`).
"""
import torch

from . import _lib
from . import ops


def eligible(blocks, x, prestrided):
    """The own-kernel path covers the configuration the C4 head runs: bf16, channels-last 4x4 pooled input whose stride-2
    sampling already happened in the pooler (`prestrided`), 3x3 / stride 1 / pad 1 / ungrouped conv2, channel counts
    multiples of 64.  Anything else stays on the library path (layers._FrozenRes5MeanFn / forward_folded)."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[0] > 0 and x.shape[2:] == (4, 4)):
        return False
    if not x.permute(0, 2, 3, 1).is_contiguous():
        return False
    for i, b in enumerate(blocks):
        c2 = b.conv2
        if c2.kernel_size != (3, 3) or c2.stride != (1, 1) or c2.padding != (1, 1) or c2.dilation != (1, 1) or c2.groups != 1:
            return False
        if b.conv1.kernel_size != (1, 1) or b.conv3.kernel_size != (1, 1) or b.conv1.padding != (0, 0):
            return False
        s1 = b.conv1.stride
        if i == 0:
            if s1 != (1, 1) and not prestrided:
                return False
            if b.shortcut is not None and b.shortcut.stride != s1:
                return False
        elif s1 != (1, 1) or b.shortcut is not None:
            return False
        if any(c % 64 for c in (b.conv1.in_channels, b.conv1.out_channels, b.conv3.out_channels)):
            return False
    return blocks[0].conv1.in_channels == x.shape[1]


class _BlockWeights:
    """GEMM operands of one frozen bottleneck, BN folded, bf16; forward and data-gradient forms."""

    def __init__(self, blk, x):
        (w1, b1), (w2, b2), (w3, b3), sc, _, _ = blk.folded_params(x, True)
        f = lambda t: None if t is None else t.float().contiguous()
        co1, ci = w1.shape[0], w1.shape[1]
        co3 = w3.shape[0]
        W1 = w1.reshape(co1, ci).contiguous()
        W2 = w2.permute(0, 2, 3, 1).reshape(co1, 9 * w2.shape[1]).contiguous()          # (co, ky, kx, ci)
        W3 = w3.reshape(co3, co1).contiguous()
        self.b1, self.b2 = f(b1), f(b2)
        self.w1, self.w2 = W1, W2
        self.has_sc = sc is not None
        if self.has_sc:
            Wsc = sc[0].reshape(co3, ci)
            self.w3 = torch.cat([W3, Wsc], 1).contiguous()                               # K = [conv2 output | block input]
            self.b3 = f(b3) if sc[1] is None else f(b3) + f(sc[1])
            self.w1t = torch.cat([W1.t(), Wsc.t()], 1).contiguous()                      # (ci, co1 + co3): [g1 | g] -> gx
        else:
            self.w3, self.b3 = W3, f(b3)
            self.w1t = W1.t().contiguous()                                               # (ci, co1)
        self.w3t = W3.t().contiguous()                                                   # (co1, co3)
        # dX[y, x] = sum_{ky', kx'} dY[y + ky' - 1, x + kx' - 1] W[:, :, 2 - ky', 2 - kx']^T : same implicit GEMM, taps flipped
        self.w2t = w2.flip(2, 3).permute(1, 2, 3, 0).reshape(w2.shape[1], 9 * co1).contiguous()
        self.c_mid, self.c_out, self.c_in = co1, co3, ci


def block_weights(blocks, x):
    out = []
    for b in blocks:
        convs = [b.conv1, b.conv2, b.conv3, b.shortcut]
        key = (x.dtype, x.device) + tuple((c.weight.data_ptr(), c.weight._version) for c in convs if c is not None)
        hit = getattr(b, "_own_gemm", None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, _BlockWeights(b, x))
            b._own_gemm = hit
        out.append(hit[1])
    return out


def _bits(M, N, dev):
    return torch.empty((M, N // 32), dtype=torch.int32, device=dev)


def res5_mean_forward(ws, x, want_bits):
    """x: (R, C, 4, 4) bf16 channels-last -> (pooled (R, C_out) fp32, [packed ReLU masks per block] | None)."""
    R = x.shape[0]
    M = 16 * R
    dev = x.device
    y = x.permute(0, 2, 3, 1).reshape(M, x.shape[1])
    masks = []
    pooled = torch.empty((R, ws[-1].c_out), dtype=torch.float32, device=dev)
    for i, w in enumerate(ws):
        last = i + 1 == len(ws)
        m1 = _bits(M, w.c_mid, dev) if want_bits else None
        m2 = _bits(M, w.c_mid, dev) if want_bits else None
        my = _bits(M, w.c_out, dev) if want_bits else None
        o1 = ops.gemm2(y, w.w1, bias=w.b1, relu=True, bits_out=m1)
        o2 = ops.gemm2(o1, w.w2, conv_c=w.c_mid, bias=w.b2, relu=True, bits_out=m2)
        kw = dict(bias=w.b3, relu=True, bits_out=my)
        if last:
            kw.update(rowmean_out=pooled, want_out=False)
        if w.has_sc:
            y = ops.gemm2(o2, w.w3, a2=y, **kw)
        else:
            y = ops.gemm2(o2, w.w3, residual=y, **kw)
        masks.append((m1, m2, my))
    return pooled, (masks if want_bits else None)


def res5_mean_backward(ws, masks, gp, R):
    """gp: (R, C_out) gradient of the pooled feature -> (R, C_in, 4, 4) bf16 channels-last data gradient."""
    M = 16 * R
    dev = gp.device
    gp = gp.detach().float()
    if gp.stride(1) != 1 or gp.stride(0) % 4 or gp.data_ptr() % 16:
        gp = gp.contiguous()
    c_out = ws[-1].c_out
    g = torch.empty((M, c_out), dtype=torch.bfloat16, device=dev)
    _lib.call("b200_mean_bwd_relu_bits", gp.data_ptr(), gp.stride(0), masks[-1][2].data_ptr(), g.data_ptr(), R, 16, c_out,
              1, ops._stream())          # bit_layout 1: the GEMM epilogue's interleaved mask words
    for i in reversed(range(len(ws))):
        w = ws[i]
        m1, m2, _ = masks[i]
        g2 = ops.gemm2(g, w.w3t, mask_bits=m2)
        g1 = ops.gemm2(g2, w.w2t, conv_c=w.c_mid, mask_bits=m1)
        prev = masks[i - 1][2] if i > 0 else None          # the block input is the previous block's post-ReLU output
        if w.has_sc:
            g = ops.gemm2(g1, w.w1t, a2=g, mask_bits=prev)
        else:
            g = ops.gemm2(g1, w.w1t, residual=g, mask_bits=prev)
    return g.reshape(R, 4, 4, ws[0].c_in).permute(0, 3, 1, 2)


class _FrozenRes5MeanOwn(torch.autograd.Function):
    """Frozen res5 + spatial mean as one autograd node on the tcgen05 GEMM (no library kernels)."""

    @staticmethod
    def forward(ctx, x, ws):
        need = ctx.needs_input_grad[0]
        pooled, masks = res5_mean_forward(ws, x, need)
        ctx.ws, ctx.masks, ctx.R = ws, masks, x.shape[0]
        return pooled

    @staticmethod
    def backward(ctx, gp):
        return res5_mean_backward(ctx.ws, ctx.masks, gp, ctx.R), None


def frozen_res5_mean(blocks, x, prestrided):
    """pooled (R, C_out) fp32, or None when the own-kernel path does not cover this configuration."""
    if not eligible(blocks, x, prestrided):
        return None
    ws = block_weights(blocks, x)
    if torch.is_grad_enabled() and x.requires_grad:
        return _FrozenRes5MeanOwn.apply(x, ws)
    return res5_mean_forward(ws, x.detach(), False)[0]
