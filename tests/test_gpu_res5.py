"""Frozen res5 stage + spatial mean on the CTA-pair tcgen05 GEMM (res5_ops.py) at FULL width
(1024 -> 512 -> 2048, three bottlenecks) against the fp32 CPU oracle (`oracle.res5`, the restatement of
roi_heads.py:313-344 + :1109) — forward and data gradient — and against the library (cuDNN) path it replaces.

Bars (north_star, bf16): elementwise |got - ref| <= 2e-2 |ref| + 2e-2 max|ref| and relative L2 <= 2e-2
  * forward (pooled feature) against the fp32 oracle;
  * forward AND data gradient against the bf16-operand / fp32-accumulate restatement of the same arithmetic
    (`_emulate`: fp64 torch, weights and every stored activation / gradient rounded to bf16 where the kernels round).
The data gradient against the fp32 oracle is NOT within 2e-2 for any bf16 implementation of this stage: the gradient
passes nine ReLU gates whose pre-activations sit near zero for a random-init network, and rounding the WEIGHTS alone to
bf16 (activations and accumulation left in fp32) already moves it by 9.3e-2 in relative L2 (cosine 0.9956); measured on
this fixture: own kernels 1.05e-1 (cosine 0.9945), cuDNN bf16 path 1.10e-1 (0.9939).  That comparison is therefore held
to "no worse than the weight-rounding floor x 1.25" and cosine >= 0.99, with the numbers above written here."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def _head(seed=0):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = "Res5ROIHeads"
    torch.manual_seed(seed)
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)}).eval()
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():                      # non-trivial FrozenBN statistics, as in a trained checkpoint
        for name, buf in m.named_buffers():
            if name.endswith("norm.weight"):
                buf.copy_(1.0 + 0.2 * torch.randn(buf.shape, generator=gen))
            elif name.endswith("norm.bias"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=gen))
            elif name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=gen))
            elif name.endswith("running_var"):
                buf.copy_(1.0 + 0.3 * torch.rand(buf.shape, generator=gen))
    for p in m.res5.parameters():
        p.requires_grad_(False)
    return m


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _rb(t):
    return t.to(torch.bfloat16).to(torch.float64)


def _unpack(words, R, C):
    """packed mask words (16 R, C / 32) int32 -> (R, C, 4, 4) bool."""
    w = words.to(torch.int64) & 0xffffffff
    pos = torch.tensor([(i >> 1) | ((i & 1) << 4) for i in range(32)], dtype=torch.int64, device=w.device)   # b200_gemm2's layout
    b = ((w[:, :, None] >> pos) & 1).reshape(16 * R, C).bool()
    return b.reshape(R, 4, 4, C).permute(0, 3, 1, 2)


def _emulate(ws, x4, gp, gates=None):
    """bf16-operand / fp32-accumulate restatement of res5_ops (fp64 torch on the GPU): x4 (R, C, 4, 4) bf16 values,
    returns (pooled, dL/dx4, gate mismatches).  Rounding points = the kernels': every stored activation and every stored
    gradient.  `gates`: the ReLU gates [(o1 > 0, o2 > 0, out > 0)] per block the backward uses instead of its own —
    the kernels' packed masks, so that the data gradient is compared on the function the kernels' forward computed
    (a gate whose pre-activation is within one bf16 rounding of zero may open on one side and not on the other)."""
    import torch.nn.functional as F
    R = x4.shape[0]
    D = torch.float64
    y = x4.to(D)
    saved = []
    for i, w in enumerate(ws):
        last = i + 1 == len(ws)
        W1 = w.w1.to(D).reshape(w.c_mid, -1, 1, 1)
        W2 = w.w2.to(D).reshape(w.c_mid, 3, 3, w.c_mid).permute(0, 3, 1, 2)
        W3 = w.w3.to(D)
        o1 = _rb(F.relu(F.conv2d(y, W1, w.b1.to(D))))
        o2 = _rb(F.relu(F.conv2d(o1, W2, w.b2.to(D), padding=1)))
        o = F.conv2d(o2, W3[:, :w.c_mid].reshape(w.c_out, w.c_mid, 1, 1), w.b3.to(D))
        o = o + (F.conv2d(y, W3[:, w.c_mid:].reshape(w.c_out, -1, 1, 1)) if w.has_sc else y)
        o = F.relu(o)
        saved.append((y, o1, o2, o, W1, W2, W3))
        y = o if last else _rb(o)
    pooled = y.mean(dim=[2, 3])
    own = [(o1 > 0, o2 > 0, o > 0) for (_, o1, o2, o, _, _, _) in saved]
    flips = 0 if gates is None else sum(int((a != b).sum()) for ga, gb in zip(own, gates) for a, b in zip(ga, gb))
    gates = own if gates is None else gates
    g = _rb(gp.to(D)[:, :, None, None].expand(-1, -1, 4, 4) / 16 * gates[-1][2])
    for i in reversed(range(len(ws))):
        w = ws[i]
        _, _, _, _, W1, W2, W3 = saved[i]
        g2 = _rb(F.conv_transpose2d(g, W3[:, :w.c_mid].reshape(w.c_out, w.c_mid, 1, 1)) * gates[i][1])
        g1 = _rb(F.conv_transpose2d(g2, W2, padding=1) * gates[i][0])
        gx = F.conv_transpose2d(g1, W1)
        gx = gx + (F.conv_transpose2d(g, W3[:, w.c_mid:].reshape(w.c_out, -1, 1, 1)) if w.has_sc else g)
        if i > 0:
            gx = gx * gates[i - 1][2]
        g = _rb(gx)
    return pooled, g, flips


def _close(got, ref, tol=2e-2):
    got, ref = got.double().cpu(), ref.double().cpu()
    torch.testing.assert_close(got, ref, rtol=tol, atol=tol * float(ref.abs().max()))
    assert _rel(got, ref) <= tol, _rel(got, ref)


@pytest.mark.parametrize("R", [24, 129])
def test_res5_own_kernels_vs_oracle_forward_backward(R):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import res5_ops
    m = _head()
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(R)
    x7 = torch.relu(torch.randn(R, 1024, 7, 7, generator=gen)) * 0.5            # the pooled 7x7 map
    x7 = x7.to(torch.bfloat16).float()                                          # both sides see the same bf16-rounded input
    gp = torch.randn(R, 2048, generator=gen)
    xr = x7.clone().requires_grad_(True)
    ref = O.res5(xr, p).mean(dim=[2, 3])                                        # reference arithmetic, fp32, CPU
    ref.backward(gp)
    gref = xr.grad[:, :, ::2, ::2]                                              # only the live bins carry gradient
    assert float(xr.grad[:, :, 1::2].abs().max()) == 0.0 and float(xr.grad[:, :, :, 1::2].abs().max()) == 0.0
    m = m.cuda()
    x4 = x7[:, :, ::2, ::2].to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert res5_ops.eligible(m.res5, x4, True)
    got = res5_ops.frozen_res5_mean(m.res5, x4, True)
    assert got.dtype == torch.float32 and got.shape == (R, 2048)
    _close(got, ref.detach())
    got.backward(gp.cuda())
    # against the bf16-operand / fp32-accumulate restatement: the forward, and the data gradient through the gates the
    # kernels' forward recorded, within the bf16 bar; the gates themselves differ from the restatement's own on < 0.2 %
    ws = res5_ops.block_weights(m.res5, x4)
    _, masks = res5_ops.res5_mean_forward(ws, x4.detach(), True)
    gates = [tuple(_unpack(b, R, c) for b, c in zip(mk, (w.c_mid, w.c_mid, w.c_out))) for mk, w in zip(masks, ws)]
    pe, ge, flips = _emulate(ws, x4.detach(), gp.cuda(), gates)
    _close(got.detach(), pe)
    _close(x4.grad.float(), ge)
    assert flips <= 2e-3 * sum(g.numel() for gs in gates for g in gs), flips
    # against the fp32 oracle: see the module docstring (weight rounding alone: 9.3e-2)
    rel = _rel(x4.grad.float().cpu(), gref)
    a, b = x4.grad.double().cpu().flatten(), gref.double().flatten()
    cos = float(a @ b / (a.norm() * b.norm()))
    assert rel <= 1.25 * 9.3e-2 and cos >= 0.99, (rel, cos)
    # inference direction (no masks, no autograd): same numbers
    with torch.no_grad():
        inf = res5_ops.frozen_res5_mean(m.res5, x4.detach(), True)
    assert torch.equal(inf, got.detach())


def test_res5_own_kernels_vs_library_path_through_the_head():
    """`Res5ROIHeads._res5_mean` dispatches to the own kernels by default and to cuDNN with RES5_IMPL = "cudnn": both see
    the same frozen weights; pooled feature and gradient agree to bf16 rounding of the intermediate activations."""
    m = _head(3).cuda()
    gen = torch.Generator().manual_seed(5)
    R = 200
    x = (torch.relu(torch.randn(R, 1024, 4, 4, generator=gen)) * 0.5).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    gp = torch.randn(R, 2048, generator=gen).cuda()
    outs = []
    for impl in ("tcgen05", "cudnn"):
        m.res5_impl = impl
        xa = x.clone().requires_grad_(True)
        with torch.enable_grad():
            pooled = m._res5_mean(xa, prestrided=True)
            pooled.backward(gp)
        outs.append((pooled.detach(), xa.grad.float()))
    _close(outs[0][0], outs[1][0], 1e-2)
    # two bf16 implementations accumulating in different orders: ReLU gates whose pre-activation is within a rounding of
    # zero open on one side only (see the module docstring; measured 8.7e-2, cosine 0.996)
    a, b = outs[0][1].double().flatten(), outs[1][1].double().flatten()
    assert _rel(outs[0][1], outs[1][1]) < 0.12 and float(a @ b / (a.norm() * b.norm())) > 0.99


def test_res5_own_kernels_full_size_is_deterministic_and_matches_library():
    """BASELINE size: 4096 ROIs (65536 pixel rows).  Bitwise reproducible; agrees with the cuDNN path."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import res5_ops
    m = _head(7).cuda()
    R = 4096
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.relu(torch.randn(R, 4, 4, 1024, device="cuda", generator=gen)) * 0.5).to(torch.bfloat16).permute(0, 3, 1, 2)
    gp = torch.randn(R, 2048, device="cuda", generator=gen)
    res = []
    for _ in range(2):
        xa = x.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        pooled = res5_ops.frozen_res5_mean(m.res5, xa, True)
        pooled.backward(gp)
        res.append((pooled.detach().clone(), xa.grad.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1].view(torch.int16), res[1][1].view(torch.int16))
    m.res5_impl = "cudnn"
    xb = x.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
    pb = m._res5_mean(xb, prestrided=True)
    pb.backward(gp)
    a, b = res[0][1].double().flatten(), xb.grad.double().flatten()
    assert _rel(res[0][0], pb.detach()) < 1e-2 and _rel(res[0][1].float(), xb.grad.float()) < 0.12
    assert float(a @ b / (a.norm() * b.norm())) > 0.99
