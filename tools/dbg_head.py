import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import emulate_head as E
from fewshotobjectdetection_imporove_via_text_feature_b200 import train_ops, config, modeling
from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, ShapeSpec
from oracle.gen_golden import synth_proposals
cfg = config.get_cfg()
cfg.MODEL.ROI_HEADS.NAME = "SematicRes5ROIHeads"; cfg.MODEL.ADDITION.NAME = "clip"; cfg.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA = 0.5
torch.manual_seed(3)
m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)}).cuda().train()
with torch.no_grad():
    m.box_predictor.cls_score.weight.mul_(20.0); m.box_predictor.bbox_pred.weight.mul_(50.0)
gen = torch.Generator().manual_seed(4)
R, K = 1024, 20
b, _ = synth_proposals(R, 600, 800, gen)
gtb = b + torch.randn(R, 4, generator=gen) * 4
gtb[:, 2:] = torch.maximum(gtb[:, 2:], gtb[:, :2] + 2)
gt = torch.randint(0, K + 1, (R,), generator=gen); gt[R // 4:] = K
x0 = torch.relu(torch.randn(R, 2048, generator=gen)).cuda()
gt, props, gtb = gt.cuda(), b.cuda(), gtb.cuda()
sa, pred = m.attention.attention, m.box_predictor
kq, vp = train_ops.text_side(m.attention)
x = x0.clone().requires_grad_(True)
losses, _, acc = train_ops.fused_head_train(x, kq, vp, sa, pred, gt, props, gtb, K, m.box2box_transform.weights, 0.5, 0.0, 1, True)
losses.sum().backward()
P = dict(W1=sa.linear1[0].weight, b1=sa.linear1[0].bias, W2=sa.linear2[0].weight, b2=sa.linear2[0].bias,
         W3=sa.linear3.weight, b3=sa.linear3.bias, Wf1=sa.ffn.linear1.weight, bf1=sa.ffn.linear1.bias,
         Wf2=sa.ffn.linear2.weight, bf2=sa.ffn.linear2.bias, gamma=sa.ffn.norm3.weight, beta=sa.ffn.norm3.bias,
         Wc=pred.cls_score.weight, bc=pred.cls_score.bias, Wb=pred.bbox_pred.weight, bb=pred.bbox_pred.bias)
L, G = E.emulate(P, x0, kq, vp, gt, props, gtb, K, m.box2box_transform.weights, 0.5)
e = x.grad.double() - G["x"]
refmax = float(G["x"].abs().max())
rowmax = e.abs().max(1).values
bad = torch.nonzero(rowmax > 2e-2 * refmax).flatten()
print("refmax", refmax, "bad rows", bad.numel(), bad[:20].tolist(), "gt", gt[bad][:20].tolist())
print("row norms of ref: fg mean %.3e bg mean %.3e" % (float(G["x"][gt < K].norm(dim=1).mean()), float(G["x"][gt == K].norm(dim=1).mean())))
r = int(bad[0])
print("row", r, "|e|", float(e[r].norm()), "|ref|", float(G["x"][r].norm()), "max e", float(e[r].abs().max()), "argmax col", int(e[r].abs().argmax()))
cols = (e[r].abs() > 2e-2 * refmax).nonzero().flatten()
print("bad cols in row:", cols.numel(), cols[:16].tolist())
print("x0 at bad cols", x0[r][cols[:8]].tolist())
got = dict(x=x.grad, **{k: v.grad for k, v in P.items()})
for k, ref in G.items():
    if k in got and got[k] is not None:
        gg = got[k].double().reshape(ref.shape)
        print("%-6s rel %.2e  max/maxref %.2e" % (k, float((gg - ref).norm() / ref.norm()), float((gg - ref).abs().max() / ref.abs().max())))
# does the error of a bad row lie along Wb rows (box branch)?
Wb = E.rb(P["Wb"].detach().double())
er = e[r]
c0 = 4 * int(gt[r])
coef = torch.linalg.lstsq(Wb[c0:c0 + 4].t(), er[:, None]).solution.flatten()
res = er - Wb[c0:c0 + 4].t() @ coef
print("projection of the row error on its 4 box rows of Wb: coef", coef.tolist(), "residual/err", float(res.norm() / er.norm()), " 1/R =", 1.0 / R)
print("---- one loss at a time")
for i in range(3):
    x = x0.clone().requires_grad_(True)
    kq, vp = train_ops.text_side(m.attention)
    losses, _, acc = train_ops.fused_head_train(x, kq, vp, sa, pred, gt, props, gtb, K, m.box2box_transform.weights, 0.5, 0.0, 1, True)
    losses[i].backward()
    w = [0.0, 0.0, 0.0]; w[i] = 1.0
    L, G = E.emulate(P, x0, kq, vp, gt, props, gtb, K, m.box2box_transform.weights, 0.5, wts=tuple(w))
    e = x.grad.double() - G["x"]
    print(i, "rel", float(e.norm() / G["x"].norm()), "rows<256 rel", float(e[:256].norm() / G["x"][:256].norm()), "rows>=256 rel", float(e[256:].norm() / G["x"][256:].norm().clamp_min(1e-30)))
print("---- intermediates, cls loss only")
train_ops._DEBUG = {}
x = x0.clone().requires_grad_(True)
kq, vp = train_ops.text_side(m.attention)
losses, _, acc = train_ops.fused_head_train(x, kq, vp, sa, pred, gt, props, gtb, K, m.box2box_transform.weights, 0.5, 0.0, 1, True)
losses[0].backward()
L, G = E.emulate(P, x0, kq, vp, gt, props, gtb, K, m.box2box_transform.weights, 0.5, wts=(1.0, 0.0, 0.0))
for k, ref in G["_dbg"].items():
    got = train_ops._DEBUG[k].double()
    got = got[:, :ref.shape[1]]
    print("%-8s rel %.3e   |ref| %.3e" % (k, float((got - ref).norm() / ref.norm().clamp_min(1e-30)), float(ref.norm())))
