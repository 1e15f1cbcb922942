"""GDL + affine and PCB kernels vs the reference golden vectors / the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_gdl_affine_golden(golden):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import modeling
    g = golden("gdl")
    aff = modeling.AffineLayer(6, bias=True).cuda()
    with torch.no_grad():
        aff.weight.copy_(T(g["w"]))
        aff.bias.copy_(T(g["b"]))
    x = T(g["x"]).cuda().requires_grad_(True)
    y = modeling.decoupled_affine(x, aff, float(g["lam"]))
    assert torch.equal(y.cpu(), T(g["y"]))                              # op-for-op rounding: bit-exact
    y.backward(T(g["g"]).cuda())
    torch.testing.assert_close(x.grad.cpu(), T(g["gx"]), rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(aff.weight.grad.cpu(), T(g["gw"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(aff.bias.grad.cpu(), T(g["gb"]), rtol=1e-5, atol=1e-6)
    # the unfused spelling of the reference works too
    x2 = T(g["x"]).cuda().requires_grad_(True)
    y2 = aff(modeling.decouple_layer(x2, float(g["lam"])))
    assert torch.equal(y2, y)
    y2.backward(T(g["g"]).cuda())
    torch.testing.assert_close(x2.grad, x.grad, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("cl_out,bf16", [(True, False), (True, True), (False, True)])
def test_gdl_affine_layouts_full_size(cl_out, bf16):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(0)
    N, C, H, W = 2, 1024, 38, 50
    x = torch.relu(torch.randn(N, C, H, W, generator=gen)).cuda().requires_grad_(True)
    w = torch.randn(1, C, 1, 1, generator=gen).cuda().requires_grad_(True)
    b = torch.randn(1, C, 1, 1, generator=gen).cuda().requires_grad_(True)
    lam = 0.01
    y = ops.gdl_affine(x, w, b, lam, torch.bfloat16 if bf16 else None, cl_out)
    ref = x.detach() * w.detach() + b.detach()
    assert y.is_contiguous(memory_format=torch.channels_last if cl_out else torch.contiguous_format)
    tol = dict(rtol=1e-2, atol=1e-2) if bf16 else dict(rtol=0, atol=0)
    torch.testing.assert_close(y.float().contiguous(), ref, **tol)
    g = torch.randn(N, C, H, W, generator=gen).cuda()
    y.backward(g.to(y.dtype))
    gq = g.to(y.dtype).float()
    torch.testing.assert_close(x.grad, gq * w.detach() * lam, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(w.grad.flatten(), (gq * x.detach()).sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(b.grad.flatten(), gq.sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)


def test_pcb_golden(golden):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, evaluation
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances
    g = golden("pcb")
    cfg = config.get_cfg()
    cfg.DATASETS.TEST = ("voc_2007_test_all1",)
    fc = torch.nn.Linear(64, g["fc_w"].shape[0]).cuda()
    with torch.no_grad():
        fc.weight.copy_(T(g["fc_w"]))
        fc.bias.copy_(T(g["fc_b"]))
    pcb = evaluation.PrototypicalCalibrationBlock(cfg, fc=fc, prototypes=T(g["protos"]))
    assert pcb.exclude_cls == list(g["exclude"])
    inst = Instances((416, 608))
    inst.pred_boxes = Boxes(T(g["boxes"]).cuda())
    inst.scores = T(g["scores_in"]).cuda()
    inst.pred_classes = T(g["classes"]).cuda()
    with torch.no_grad():
        feats = pcb.extract_roi_features(T(g["conv_feature"]), [inst.pred_boxes])
        torch.testing.assert_close(feats.cpu(), T(g["feats_all"]), rtol=1e-4, atol=1e-4)
        dts = pcb.execute_calibration([{"conv_feature": T(g["conv_feature"])}], [{"instances": inst}])
    torch.testing.assert_close(dts[0]["instances"].scores.cpu(), T(g["scores_out"]), rtol=1e-5, atol=1e-5)


def test_pcb_random_vs_oracle():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    gen = torch.Generator().manual_seed(3)
    n, D, K = 100, 1000, 20
    s = torch.sort(torch.rand(n, generator=gen) * 1.05, descending=True).values
    s[-10:] *= 0.04
    f, p = torch.randn(n, D, generator=gen), torch.randn(K, D, generator=gen)
    c = torch.randint(0, K, (n,), generator=gen)
    excl = set(range(15))
    il, ir = int((s > 1.0).sum()), int((s > 0.05).sum())
    ref = O.pcb_calibrate(s, f[il:ir], p, c, 0.5, excl)
    mask = torch.zeros(K, dtype=torch.uint8)
    mask[list(excl)] = 1
    out = ops.pcb_cosine_blend_(s.clone().cuda(), f.cuda(), p.cuda(), c.cuda(), mask.cuda(), 0.5, 0.05, 1.0)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)


def test_pcb_one_argument_construction_end_to_end():
    """The reference's call pattern (evaluator.py:88-110): PrototypicalCalibrationBlock(cfg) builds the ImageNet extractor
    itself, prototypes come from the reference's dataset dicts (image file + gt Instances in the resized frame), then
    execute_calibration(inputs, dts) on file-name inputs.  Checked against the oracle's restatement of the blend fed with
    the block's own ROI features."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, evaluation
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances
    cfg = config.get_cfg()
    cfg.MODEL.DEVICE = "cuda"
    cfg.DATASETS.TEST = ("voc_2007_test_novel1",)             # no excluded classes
    rs = np.random.RandomState(5)
    images = {"img%d.jpg" % i: rs.randint(0, 256, (160, 224, 3)).astype(np.uint8) for i in range(4)}
    torch.manual_seed(0)
    pcb = evaluation.PrototypicalCalibrationBlock(cfg, image_reader=lambda f: images[f])
    assert pcb.imagenet_model is not None and not pcb.imagenet_model.training
    support = []
    for i in range(3):
        inst = Instances((80, 112))                           # annotations live in a frame half the file's size
        inst.gt_boxes = Boxes(torch.tensor([[4.0, 6.0, 60.0, 70.0], [30.0, 10.0, 100.0, 64.0]]))
        inst.gt_classes = torch.tensor([i % 2, 2])
        support.append({"file_name": "img%d.jpg" % i, "instances": inst})
    protos = pcb.build_prototypes(support)
    assert sorted(protos) == [0, 1, 2] and protos[2].shape == (1, 1000)
    det = Instances((160, 224))
    gen = torch.Generator().manual_seed(1)
    n = 12
    b = torch.rand(n, 4, generator=gen) * 80
    b[:, 2:] += torch.tensor([100.0, 60.0])
    det.pred_boxes = Boxes(b.cuda())
    s = torch.sort(torch.rand(n, generator=gen), descending=True).values
    s[0], s[-1] = 1.2, 0.01                                   # one above the upper, one below the lower threshold
    det.scores = s.clone().cuda()
    det.pred_classes = torch.randint(0, 3, (n,), generator=gen).cuda()
    feats = pcb.extract_roi_features(images["img3.jpg"], [det.pred_boxes]).float().cpu()
    out = pcb.execute_calibration([{"file_name": "img3.jpg"}], [{"instances": det}])
    il, ir = int((s > cfg.TEST.PCB_UPPER).sum()), int((s > cfg.TEST.PCB_LOWER).sum())
    pm = torch.cat([protos[c] for c in range(3)], 0)
    ref = O.pcb_calibrate(s, feats[il:ir], pm, det.pred_classes.cpu(), cfg.TEST.PCB_ALPHA, set())
    torch.testing.assert_close(out[0]["instances"].scores.cpu(), ref, rtol=1e-4, atol=1e-5)
    assert float(out[0]["instances"].scores[0]) == pytest.approx(1.2) and float(out[0]["instances"].scores[-1]) == pytest.approx(0.01)
