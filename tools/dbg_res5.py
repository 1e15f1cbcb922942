import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_gpu_res5 as T
from fewshotobjectdetection_imporove_via_text_feature_b200 import res5_ops, ops, _lib
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
m = T._head().cuda()
R = 24
gen = torch.Generator().manual_seed(R)
x7 = (torch.relu(torch.randn(R, 1024, 7, 7, generator=gen)) * 0.5).to(torch.bfloat16).float()
gp = torch.randn(R, 2048, generator=gen).cuda()
x4 = x7[:, :, ::2, ::2].to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
ws = res5_ops.block_weights(m.res5, x4)
D = torch.float64
rb = T._rb
def nhwc(t):  # (R,C,4,4) fp64 -> (M,C)
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
def unpack(words, N):
    w = words.to(torch.int64) & 0xffffffff
    pos = torch.tensor([(i >> 1) | ((i & 1) << 4) for i in range(32)], dtype=torch.int64, device=w.device)
    return ((w[:, :, None] >> pos) & 1).reshape(w.shape[0], -1)[:, :N].bool()
def cmp(name, got, ref):
    got, ref = got.double(), ref.double()
    d = (got - ref).abs()
    print("%-14s rel %.3e  maxabs %.3e refmax %.3e  frac>2e-2max %.4f" % (name, float((got - ref).norm() / ref.norm()), float(d.max()), float(ref.abs().max()), float((d > 2e-2 * ref.abs().max()).float().mean())))
M = 16 * R
y = x4.permute(0, 2, 3, 1).reshape(M, 1024)
ye = x4.to(D)
saved = []
for i, w in enumerate(ws):
    m1 = torch.empty(M, w.c_mid // 32, dtype=torch.int32, device="cuda"); m2 = torch.empty_like(m1)
    my = torch.empty(M, w.c_out // 32, dtype=torch.int32, device="cuda")
    o1 = ops.gemm2(y, w.w1, bias=w.b1, relu=True, bits_out=m1)
    o2 = ops.gemm2(o1, w.w2, conv_c=w.c_mid, bias=w.b2, relu=True, bits_out=m2)
    yn = ops.gemm2(o2, w.w3, a2=y, bias=w.b3, relu=True, bits_out=my) if w.has_sc else ops.gemm2(o2, w.w3, residual=y, bias=w.b3, relu=True, bits_out=my)
    W1 = w.w1.to(D).reshape(w.c_mid, -1, 1, 1); W2 = w.w2.to(D).reshape(w.c_mid, 3, 3, w.c_mid).permute(0, 3, 1, 2); W3 = w.w3.to(D)
    e1 = rb(F.relu(F.conv2d(ye, W1, w.b1.to(D))))
    e2 = rb(F.relu(F.conv2d(e1, W2, w.b2.to(D), padding=1)))
    eo = F.conv2d(e2, W3[:, :w.c_mid].reshape(w.c_out, w.c_mid, 1, 1), w.b3.to(D)) + (F.conv2d(ye, W3[:, w.c_mid:].reshape(w.c_out, -1, 1, 1)) if w.has_sc else ye)
    eo = rb(F.relu(eo))
    cmp("b%d o1" % i, o1, nhwc(e1)); cmp("b%d o2" % i, o2, nhwc(e2)); cmp("b%d y" % i, yn, nhwc(eo))
    print("   mask mismatches", int((unpack(m1, w.c_mid) != (nhwc(e1) > 0)).sum()), int((unpack(m2, w.c_mid) != (nhwc(e2) > 0)).sum()), int((unpack(my, w.c_out) != (nhwc(eo) > 0)).sum()),
          " own-consistency", int((unpack(m1, w.c_mid) != (o1 > 0)).sum()), int((unpack(my, w.c_out) != (yn > 0)).sum()))
    saved.append((y, ye, (m1, m2, my), (e1, e2, eo), (W1, W2, W3)))
    y, ye = yn, eo
# backward with EMULATION masks on both sides: isolates the backward GEMMs
g = torch.empty(M, 2048, dtype=torch.bfloat16, device="cuda")
gpc = gp.contiguous()
_lib.call("b200_mean_bwd_relu_bits", gpc.data_ptr(), gpc.stride(0), saved[-1][2][2].data_ptr(), g.data_ptr(), R, 16, 2048, ops._stream())
ge = rb(gp.to(D)[:, :, None, None].expand(-1, -1, 4, 4) / 16 * (ye > 0))
cmp("g3", g, nhwc(ge))
for i in reversed(range(3)):
    w = ws[i]
    yin, yine, (m1, m2, my), (e1, e2, eo), (W1, W2, W3) = saved[i]
    g2 = ops.gemm2(g, w.w3t, mask_bits=m2)
    g1 = ops.gemm2(g2, w.w2t, conv_c=w.c_mid, mask_bits=m1)
    prev = saved[i - 1][2][2] if i > 0 else None
    gx = ops.gemm2(g1, w.w1t, a2=g, mask_bits=prev) if w.has_sc else ops.gemm2(g1, w.w1t, residual=g, mask_bits=prev)
    eg2 = rb(F.conv_transpose2d(ge, W3[:, :w.c_mid].reshape(w.c_out, w.c_mid, 1, 1)) * (e2 > 0))
    eg1 = rb(F.conv_transpose2d(eg2, W2, padding=1) * (e1 > 0))
    egx = F.conv_transpose2d(eg1, W1) + (F.conv_transpose2d(ge, W3[:, w.c_mid:].reshape(w.c_out, -1, 1, 1)) if w.has_sc else ge)
    if i > 0:
        egx = egx * (yine > 0)
    egx = rb(egx)
    cmp("b%d g2" % i, g2, nhwc(eg2)); cmp("b%d g1" % i, g1, nhwc(eg1)); cmp("b%d gx" % i, gx, nhwc(egx))
    # same step but fed with the emulation's input gradient: per-GEMM error only
    gin = nhwc(ge).to(torch.bfloat16).contiguous()
    g2b = ops.gemm2(gin, w.w3t, mask_bits=m2)
    cmp("  g2|exact-in", g2b, nhwc(eg2))
    g1b = ops.gemm2(nhwc(eg2).to(torch.bfloat16).contiguous(), w.w2t, conv_c=w.c_mid, mask_bits=m1)
    cmp("  g1|exact-in", g1b, nhwc(eg1))
    g, ge = gx, egx
