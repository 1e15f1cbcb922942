import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train import _build, _proposals, _rel
g = np.load("tests/golden/train_step.npz")
m = _build(g)
x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
props = _proposals(g)
losses, logits = m.fused_train_losses(x, props, props[0].gt_classes)
print({k: (float(v), float(g["loss." + k])) for k, v in losses.items()})
for sel in (None, "loss_cls", "loss_box_reg", "loss_attentive"):
    m.zero_grad(set_to_none=True); x.grad = None
    losses, logits = m.fused_train_losses(x, props, props[0].gt_classes)
    (sum(losses.values()) if sel is None else losses[sel]).backward()
    gx = x.grad.clone()
    # torch path on the same device
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads.fast_rcnn import FastRCNNOutputs
    fused = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    x2 = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    att_out, att_loss = m.forward_att(x2, props[0].gt_classes)
    o = FastRCNNOutputs(m.box2box_transform, att_out["pred_logits"], att_out["pred_bbox"], props, m.smooth_l1_beta)
    L = dict(o.losses()); L.update(att_loss)
    (sum(L.values()) if sel is None else L[sel]).backward()
    print("==", sel, "grad_x fused vs torch:", _rel(gx, x2.grad), " |gx|", float(x2.grad.norm()))
    if sel is None:
        print("   grad_x fused vs golden:", _rel(gx.cpu(), torch.from_numpy(g["grad_x"])), " torch vs golden", _rel(x2.grad.cpu(), torch.from_numpy(g["grad_x"])))
    for n, p in m.named_parameters():
        if p.grad is not None and n in fused:
            print("   %-45s %.4f" % (n, _rel(fused[n], p.grad)))
