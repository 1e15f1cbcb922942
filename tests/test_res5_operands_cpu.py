"""CPU pinning of the GEMM operand preparation of res5 on the own kernels (res5_ops._BlockWeights): BatchNorm folding, the
(co, ky, kx, ci) layout of the implicit 3x3 GEMM, the K-concatenated [conv3 | shortcut] operand and the transposed / tap-flipped
operands of the data gradient — each used in a plain fp32 torch restatement of the GEMM formulation and compared with the block's
own convolutions / autograd.  No GPU: the kernels that consume these operands are tested in tests/test_gpu_gemm2.py /
tests/test_gpu_res5.py."""
import torch
import torch.nn.functional as F


def _block(cin, cout, mid, with_bn_stats=True):
    from fewshotobjectdetection_imporove_via_text_feature_b200.layers import BottleneckBlock
    torch.manual_seed(cin + cout)
    b = BottleneckBlock(cin, cout, bottleneck_channels=mid, stride=1, norm="FrozenBN", stride_in_1x1=True).eval()
    if with_bn_stats:
        for c in (b.conv1, b.conv2, b.conv3, b.shortcut):
            if c is not None and getattr(c, "norm", None) is not None:
                n = c.norm
                n.weight.copy_(torch.rand_like(n.weight) + 0.5)
                n.bias.copy_(torch.randn_like(n.bias) * 0.1)
                n.running_mean.copy_(torch.randn_like(n.running_mean) * 0.1)
                n.running_var.copy_(torch.rand_like(n.running_var) + 0.5)
    return b


def _im2col(y, taps_flipped=False):
    """(R, 4, 4, c) -> (R * 16, 9 c): tap-major (ky, kx), channel-minor, zero padding 1 — the A operand the 4-D TMA box fetches."""
    R, H, W, c = y.shape
    p = F.pad(y, (0, 0, 1, 1, 1, 1))
    cols = [p[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)]
    return torch.cat(cols, dim=-1).reshape(R * H * W, 9 * c)


def _forward_gemm_form(w, x_rows, R):
    y1 = torch.relu(x_rows @ w.w1.float().t() + w.b1)
    y2 = torch.relu(_im2col(y1.reshape(R, 4, 4, -1)) @ w.w2.float().t() + w.b2)
    a3 = torch.cat([y2, x_rows], 1) if w.has_sc else y2
    y3 = a3 @ w.w3.float().t() + w.b3
    if not w.has_sc:
        y3 = y3 + x_rows
    return y1, y2, torch.relu(y3)


def test_block_operands_reproduce_the_convolutions_and_their_data_gradient():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import res5_ops
    R = 5
    for cin, cout, mid in ((64, 128, 64), (128, 128, 64)):          # with shortcut (block 0) / identity residual (blocks 1, 2)
        blk = _block(cin, cout, mid)
        x = torch.relu(torch.randn(R, cin, 4, 4)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        with torch.no_grad():
            w = res5_ops._BlockWeights(blk, x.detach())
        assert w.has_sc == (cin != cout) and (w.c_in, w.c_mid, w.c_out) == (cin, mid, cout)
        ref = blk(x)
        x_rows = x.detach().permute(0, 2, 3, 1).reshape(R * 16, cin)
        y1, y2, y3 = _forward_gemm_form(w, x_rows, R)
        torch.testing.assert_close(y3.reshape(R, 4, 4, cout).permute(0, 3, 1, 2), ref.detach(), rtol=1e-4, atol=1e-4)
        # data gradient in GEMM form: gate, conv3^T, gate, flipped-tap implicit GEMM, gate, [conv1^T | shortcut^T] (+ identity fan-in)
        g = torch.randn_like(ref)
        ref.backward(g)
        g_rows = g.permute(0, 2, 3, 1).reshape(R * 16, cout) * (y3 > 0)
        g2 = (g_rows @ w.w3t.float().t()) * (y2 > 0)                                   # (R16, mid)
        g1 = (_im2col(g2.reshape(R, 4, 4, mid)) @ w.w2t.float().t()) * (y1 > 0)
        if w.has_sc:
            gx = torch.cat([g1, g_rows], 1) @ w.w1t.float().t()
        else:
            gx = g1 @ w.w1t.float().t() + g_rows
        torch.testing.assert_close(gx.reshape(R, 4, 4, cin).permute(0, 3, 1, 2), x.grad, rtol=1e-4, atol=1e-4)


def test_eligibility_rules():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import res5_ops
    blk = _block(64, 128, 64)
    x = torch.zeros(2, 64, 4, 4, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert not res5_ops.eligible([blk], x, True)                     # CPU tensor: the own kernels need the GPU
