"""bf16-operand / fp32-accumulate restatement of the fused fine-tune node (train_ops._FusedHeadTrain) in fp64 torch.

Same arithmetic as the reference's modules (attentive_modules.py:114-177, fast_rcnn.py:403-417 + :222-304,
roi_heads.py:1077-1081) and their autograd, with a bf16 rounding (`rb`) exactly where the kernels store a bf16 tensor:
GEMM operands (weights, activations, incoming gradients) and the bf16 outputs of the elementwise kernels.  Everything else
is carried in fp64, so what remains between this and the kernels is fp32 accumulation order.
TEST INFRASTRUCTURE (oracle/): imported by tests/, __graft_entry__.smoke() and tools/ only, never by the product."""
import torch

D = torch.float64


def rb(t):
    return t.to(torch.bfloat16).to(D)


def box_deltas(src, dst, w):
    sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
    sx, sy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
    tw, th = dst[:, 2] - dst[:, 0], dst[:, 3] - dst[:, 1]
    tx, ty = dst[:, 0] + 0.5 * tw, dst[:, 1] + 0.5 * th
    return torch.stack((w[0] * (tx - sx) / sw, w[1] * (ty - sy) / sh, w[2] * torch.log(tw / sw), w[3] * torch.log(th / sh)), 1)


def emulate(P, x, kq, vp, gt, props, gtb, K, box_w=(10.0, 10.0, 5.0, 5.0), beta=0.0, eps=1e-5, wts=(1.0, 1.0, 1.0)):
    """P: dict of fp32 parameters (W1,b1,W2,b2,W3,b3,Wf1,bf1,Wf2,bf2,gamma,beta,Wc,bc,Wb,bb); x (R,d) fp32 pooled feature;
    kq (L,d), vp (L,d) the text-side operands; gt (R,) int64; props / gtb (R,4).  No dropout.
    P with "Wo", "bo" and "T" (text prototypes (K+1, D)): the CrossOutput classifier, logits = relu(zd Wo^T + bo) T^T
    (roi_heads.py:1154-1171) instead of cls_score, and no attentive loss.
    Returns (losses dict, grads dict) for sum(losses)."""
    f = lambda t: t.detach().to(D)
    cross = "Wo" in P
    W = {k: rb(f(v)) for k, v in P.items() if k.startswith("W") or k == "T"}   # bf16 GEMM operands
    b = {k: f(v) for k, v in P.items() if not (k.startswith("W") or k == "T")}
    if cross:
        wts = (wts[0], wts[1], 0.0)
    x, vp = f(x), f(vp)
    kqb = rb(f(kq))
    R, d = x.shape
    h = d // 2
    L = kq.shape[0]
    xb = rb(x)
    S = xb @ kqb.t()
    attn = torch.softmax(S, 1)
    O = attn @ vp
    p1, p2 = rb(O * x), rb(x - O)
    o1 = rb(torch.relu(p1 @ W["W1"].t() + b["b1"]))
    o2 = rb(torch.relu(p2 @ W["W2"].t() + b["b2"]))
    xcat = torch.cat([o1, o2, xb], 1)
    y = xcat @ W["W3"].t() + b["b3"]
    yb = rb(y)
    hdn = rb(torch.relu(yb @ W["Wf1"].t() + b["bf1"]))
    y2 = hdn @ W["Wf2"].t() + b["bf2"]
    u = y + y2
    mu, var = u.mean(1, keepdim=True), u.var(1, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    xh = (u - mu) * rstd
    zpre = xh * b["gamma"] + b["beta"]
    z = torch.relu(zpre)
    zd = rb(z)
    if cross:
        av = rb(torch.relu(zd @ W["Wo"].t() + b["bo"]))
        logits = av @ W["T"].t()
    else:
        logits = zd @ W["Wc"].t() + b["bc"]
    deltas = xb @ W["Wb"].t() + b["bb"]
    C1 = K + 1
    gt = gt.to(torch.int64)
    fg = (gt >= 0) & (gt < K)
    lse = torch.logsumexp(logits, 1)
    ar = torch.arange(R, device=x.device)
    loss_cls = (lse - logits[ar, gt]).mean()
    loss_att = (torch.logsumexp(attn, 1) - attn[ar, gt]).mean()
    tgt = box_deltas(f(props), f(gtb), box_w)
    cols = 4 * gt.clamp(max=K - 1)[:, None] + torch.arange(4, device=x.device)
    diff = (deltas.gather(1, cols) - tgt) * fg[:, None]
    n = diff.abs()
    lb = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n * n / beta, n - 0.5 * beta)
    loss_box = (lb * fg[:, None]).sum() / R
    losses = {"loss_cls": loss_cls, "loss_box_reg": loss_box, "loss_attentive": loss_att}
    if cross:
        losses.pop("loss_attentive")
    # ---- backward of sum(losses) -----------------------------------------------------------------------------------
    onehot = torch.zeros(R, C1, dtype=D, device=x.device)
    onehot[ar, gt] = 1
    dlogits = rb(wts[0] * (torch.softmax(logits, 1) - onehot) / R)
    gb = torch.sign(diff) if beta < 1e-5 else torch.where(n < beta, diff / beta, torch.sign(diff))
    ddeltas = torch.zeros_like(deltas)
    ddeltas.scatter_(1, cols, wts[1] * gb * fg[:, None] / R)
    ddeltas = rb(ddeltas)
    oh_att = torch.zeros(R, L, dtype=D, device=x.device)
    oh_att[ar, gt] = 1
    dattn_ext = wts[2] * (torch.softmax(attn, 1) - oh_att) / R
    G = {}
    if cross:
        da = rb((dlogits @ W["T"]) * (av > 0))
        G["Wo"], G["bo"] = da.t() @ zd, da.sum(0)
        dzd = rb(da @ W["Wo"])
    else:
        G["Wc"], G["bc"] = dlogits.t() @ zd, dlogits.sum(0)
        dzd = rb(dlogits @ W["Wc"])
    G["Wb"], G["bb"] = ddeltas.t() @ xb, ddeltas.sum(0)
    dx = ddeltas @ W["Wb"]
    dz = dzd * (zpre > 0)
    G["gamma"], G["beta"] = (dz * xh).sum(0), dz.sum(0)
    dxh = dz * b["gamma"]
    du = rstd * (dxh - dxh.mean(1, keepdim=True) - xh * (dxh * xh).mean(1, keepdim=True))
    dub = rb(du)
    G["Wf2"], G["bf2"] = dub.t() @ hdn, dub.sum(0)
    dhdn = rb((dub @ W["Wf2"]) * (hdn > 0))
    G["Wf1"], G["bf1"] = dhdn.t() @ yb, dhdn.sum(0)
    dy = du + dhdn @ W["Wf1"]
    dyb = rb(dy)
    G["W3"], G["b3"] = dyb.t() @ xcat, dy.sum(0)
    do12 = rb((dyb @ W["W3"][:, :d]) * (xcat[:, :d] > 0))
    t_l3 = dyb @ W["W3"][:, d:]
    dx = dx + t_l3
    do1, do2 = do12[:, :h], do12[:, h:]
    G["W1"], G["b1"] = do1.t() @ p1, do1.sum(0)
    G["W2"], G["b2"] = do2.t() @ p2, do2.sum(0)
    dp1, dp2 = rb(do1 @ W["W1"]), rb(do2 @ W["W2"])
    t_att = dp1 * O + dp2
    dx = dx + t_att
    dO = dp1 * x - dp2
    dA = dO @ vp.t() + dattn_ext
    dS = attn * (dA - (attn * dA).sum(1, keepdim=True))
    dOb, dSb = rb(dO), rb(dS)
    G["vp"] = rb(attn).t() @ dOb
    G["kq"] = dSb.t() @ xb
    t_s = dSb @ kqb
    G["x"] = dx + t_s
    # dL/dx is the sum of four branch gradients that largely cancel (direct path through linear3 against the attention
    # path): the scale that bf16-level differences of the branches are relative to is the largest branch, not the sum
    G["_x_branch_max"] = max(float(t.abs().max()) for t in (t_l3, t_att, t_s, G["x"]))
    G["_x_branch_norms"] = [float(t.norm()) for t in (t_l3, t_att, t_s, G["x"])]
    G["_dbg"] = dict(dx_box=ddeltas @ W["Wb"], dx_l3=ddeltas @ W["Wb"] + t_l3, dx_att=ddeltas @ W["Wb"] + t_l3 + t_att, dzd=dzd,
                     dlogits=dlogits, dyb=dyb, do12=do12, dub=dub, dhdn=dhdn, du=dy, dp1=dp1, dp2=dp2, dO=dOb, dS=dSb)
    return losses, G
