"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200roi.h declares (no compute
calls), the product refuses to run without CUDA (no CPU fallback), and the host-side mirror of the reference API
(registries, state-dict names, containers, class lists) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _build, _lib
    path = _build.build()
    L = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "b200roi.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", header))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L.b200_abi_version.restype = ctypes.c_int
    assert L.b200_abi_version() == 3
    # every declaration cites the reference interface it replaces
    assert header.count("defrcn/") >= 8


def test_no_cpu_fallback():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, modeling, ops
    with pytest.raises(_lib.B200Error):
        ops.roi_align(torch.zeros(1, 8, 4, 4), torch.zeros(1, 5), 7, 1 / 16)
    with pytest.raises(RuntimeError):
        modeling.AffineLayer(8, bias=True)(torch.zeros(1, 8, 4, 4))
    # the oracle is never imported by the product
    import sys
    pkg = "fewshotobjectdetection_imporove_via_text_feature_b200"
    for root, _, files in os.walk(os.path.join(ROOT, pkg)):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_registries_and_state_dict_names():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import ShapeSpec
    for name in ("Res5ROIHeads", "SematicRes5ROIHeads", "SematicRes5ROIHeadsCrossOutput"):
        assert modeling.ROI_HEADS_REGISTRY.get(name).__name__ == name
    for name in ("FastRCNNOutputLayers", "FastRCNNAttentionOutputLayers"):
        assert modeling.ROI_HEADS_OUTPUT_REGISTRY.get(name).__name__ == name
    with pytest.raises(KeyError):
        modeling.ROI_HEADS_REGISTRY.get("NoSuchHead")
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME, cfg.MODEL.ADDITION.NAME = "SematicRes5ROIHeads", "clip"
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 4, 1
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=16, stride=16)})
    keys = set(m.state_dict().keys())
    # names a reference checkpoint / tools/model_surgery.py expects
    for k in ["res5.0.conv1.weight", "res5.0.conv1.norm.running_var", "res5.0.shortcut.weight", "res5.2.conv3.norm.bias",
              "box_predictor.cls_score.weight", "box_predictor.bbox_pred.bias", "attention.query_projection.weight",
              "attention.output_projection.bias", "attention.key_projection.weight", "attention.value_projection.bias",
              "attention.attention.dummy", "attention.attention.w_q.weight", "attention.attention.w_k.weight",
              "attention.attention.w_v.weight", "attention.attention.linear1.0.weight", "attention.attention.linear2.0.bias",
              "attention.attention.linear3.weight", "attention.attention.ffn.linear1.weight",
              "attention.attention.ffn.linear2.bias", "attention.attention.ffn.norm3.weight", "output_projection.weight",
              "sematic_projection.bias", "projection_matrix"]:
        assert k in keys, k
    assert not any("embed" in k or "bg_feature" in k for k in keys)       # text tensors are not checkpointed
    assert m.attention.embed.shape == (20, 512) and m.attention.bg_feature.shape == (1, 512)
    assert m.box_predictor.cls_score.out_features == 21 and m.box_predictor.bbox_pred.out_features == 80


def test_golden_state_dict_loads(golden):
    """The reference's own state dict (from the golden fixture) loads with strict=True."""
    import numpy as np
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import ShapeSpec
    g = golden("head_tiny")
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME, cfg.MODEL.ADDITION.NAME = "SematicRes5ROIHeads", "clip"
    cfg.MODEL.RESNETS.RES2_OUT_CHANNELS, cfg.MODEL.RESNETS.WIDTH_PER_GROUP = 4, 1
    m = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=16, stride=16)})
    ours = set(m.state_dict().keys())
    ref = {k for k in g.files if "." in k or k == "projection_matrix"}
    assert ours == ref, ours ^ ref
    m.load_state_dict({k: torch.from_numpy(np.asarray(g[k])) for k in ours}, strict=True)


def test_structures_and_class_names():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances, pairwise_iou
    from fewshotobjectdetection_imporove_via_text_feature_b200.utils import class_embedding as ce
    b = Boxes(torch.tensor([[0.0, 0.0, 10.0, 10.0], [-5.0, 2.0, 30.0, 50.0]]))
    b.clip((20, 25))
    assert b.tensor.tolist() == [[0, 0, 10, 10], [0, 2, 25, 20]]
    inst = Instances((20, 25), pred_boxes=b, scores=torch.tensor([0.9, 0.1]))
    assert len(inst) == 2 and inst.image_size == (20, 25) and inst.has("scores") and len(inst[torch.tensor([1])]) == 1
    with pytest.raises(AttributeError):
        inst.nothing
    iou = pairwise_iou(b, b)
    assert torch.allclose(iou.diag(), torch.ones(2))
    cfg = config.get_cfg()
    cfg.DATASETS.TRAIN = ("voc_2007_trainval_novel1_10shot_seed0",)
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = 5
    assert ce.get_class_name(cfg) == ["bird", "bus", "cow", "motorbike", "sofa"]
    cfg.DATASETS.TRAIN = ("voc_2007_trainval_all1_1shot_seed0",)
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = 20
    names = ce.get_class_name(cfg)
    assert names[:2] == ["aeroplane", "bicycle"] and names[15:] == ["bird", "bus", "cow", "motorbike", "sofa"]
    e = ce.get_class_embed(names, "clip")
    assert e.shape == (20, 512) and torch.allclose(e.norm(dim=1), torch.ones(20), atol=1e-5)
    bg = ce.create_normalized_orthogonal_tensor(e.mean(0, keepdim=True), torch.Generator().manual_seed(0))
    assert abs(float(bg.norm()) - 1) < 1e-5
    cfg.merge_from_list(["MODEL.ROI_HEADS.NMS_THRESH_TEST", 0.3])
    assert cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST == 0.3
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MODEL.ROI_HEADS.TEACHER_TRAINING", True])   # stale key of the reference's scripts


def test_matcher_and_sampler():
    from fewshotobjectdetection_imporove_via_text_feature_b200.layers import Matcher, subsample_labels
    iou = torch.tensor([[0.7, 0.2, 0.4], [0.1, 0.6, 0.45]])
    idx, lab = Matcher([0.5], [0, 1])(iou)
    assert idx.tolist() == [0, 1, 1] and lab.tolist() == [1, 1, 0]
    torch.manual_seed(0)
    labels = torch.tensor([0, 1, 20, 20, 20, -1, 3, 20])
    fg, bg = subsample_labels(labels, 4, 0.25, 20)
    assert len(fg) == 1 and len(bg) == 3 and all(labels[i] == 20 for i in bg) and labels[fg[0]] in (0, 1, 3)


def test_losses_match_golden(golden):
    """FastRCNNOutputs.losses on CPU tensors (pure torch, no kernels involved) vs the reference's numbers."""
    import numpy as np
    from fewshotobjectdetection_imporove_via_text_feature_b200.layers import Box2BoxTransform
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling import FastRCNNOutputs
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances
    g = golden("losses")
    T = lambda k: torch.from_numpy(np.asarray(g[k]))
    inst = Instances((600, 800), proposal_boxes=Boxes(T("props")), gt_boxes=Boxes(T("gt_boxes")), gt_classes=T("gt_classes"))
    o = FastRCNNOutputs(Box2BoxTransform((10.0, 10.0, 5.0, 5.0)), T("logits"), T("deltas"), [inst], 0.0)
    L = o.losses()
    assert abs(float(L["loss_cls"]) - float(g["loss_cls"])) < 1e-6
    assert abs(float(L["loss_box_reg"]) - float(g["loss_box_reg"])) < 1e-6


def test_kd_losses_match_golden(golden):
    import numpy as np
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads import loss_fn_kd, loss_fn_kd_only
    g = golden("teacher")
    T = lambda k: torch.from_numpy(np.asarray(g[k]))
    params = {"alpha": float(g["alpha"]), "temperature": float(g["T"])}
    K = g["s_out"].shape[1] - 1
    assert abs(float(loss_fn_kd(T("s_out"), T("labels"), T("t_out"), params)) - float(g["kd"])) < 1e-6
    assert abs(float(loss_fn_kd_only(T("s_out"), T("labels"), K, T("t_out"), params)) - float(g["kd_only"])) < 1e-6


@pytest.mark.parametrize("tag,name", [("lv", "LV_attention"), ("td", "LV_attention_textDomination")])
def test_teacher_attention_matches_reference(golden, tag, name):
    """Teacher attentions are differentiable torch modules (they need GT labels and a backward), so their parity is
    checked on the CPU against the reference's own forward; weights are replayed by seed, which also pins the
    state-dict names and shapes."""
    import numpy as np
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling import roi_heads as RH
    from oracle.gen_golden import seeded_fill
    g = golden("teacher")
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    m = getattr(RH, name)(32, cfg=cfg, class_embed=torch.from_numpy(g[tag + "_embed"])).eval()
    keys = seeded_fill(m, 77)
    assert ["%s:%s" % (k, "x".join(map(str, s))) for k, s in keys] == list(g[tag + "_keys"])
    with torch.no_grad():
        _, out = m(torch.from_numpy(g["x"]), torch.from_numpy(g["lab"]))
    torch.testing.assert_close(out["sim2stext"], torch.from_numpy(g[tag + "_sim2stext"]), rtol=1e-4, atol=1e-5)


def test_teacher_vkv_variants_run_and_differ():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling import roi_heads as RH
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    torch.manual_seed(0)
    x, lab = torch.relu(torch.randn(10, 32)), torch.randint(0, 21, (10,))
    for base, vkv in (("LV_attention", "LV_attention_VKV"), ("LV_attention_textDomination", "LV_attention_textDomination_VKV")):
        torch.manual_seed(1)
        a = getattr(RH, base)(32, cfg=cfg)
        torch.manual_seed(1)
        b = getattr(RH, vkv)(32, cfg=cfg)
        b.load_state_dict(a.state_dict())
        assert a.float() is a and a.embed.dtype == torch.float32         # Module._apply reaches the embedding table too
        ya, yb = a(x, lab)[1]["sim2stext"], b(x, lab)[1]["sim2stext"]
        assert ya.shape == yb.shape == (1, 10, 32) and not torch.allclose(ya, yb)
        yb.sum().backward()
        assert b.w_bg.grad is not None and b.attention.w_q.weight.grad is not None


def test_eval_formats_match_reference_evaluator(golden):
    """VOC text lines and COCO records vs the reference's own PascalVOCDetectionEvaluator.process."""
    import json
    from fewshotobjectdetection_imporove_via_text_feature_b200.evaluation import detection_formats as F
    g = golden("eval_formats")
    det = dict(boxes=torch.from_numpy(g["boxes"]), scores=torch.from_numpy(g["scores"]),
               classes=torch.from_numpy(g["classes"]), counts=torch.from_numpy(g["counts"]).int())
    ids = [str(s) for s in g["ids"]]
    b, s, c, n = F.pack_batch(det)
    assert np.array_equal(b, g["boxes"]) and np.array_equal(s, g["scores"]) and np.array_equal(n, g["counts"])
    lines = F.voc_prediction_lines(ids, b, s, c, n)
    want = json.loads(str(g["voc_lines"]))
    assert {str(k): v for k, v in lines.items()} == want
    assert F.coco_json_records(ids, b, s, c, n) == json.loads(str(g["coco"]))


def test_bench_cli_contract(monkeypatch):
    """bench.py's flags and defaults as the driver uses them (`--gpus N --steps K --warmup W [--impl reference]`), and the
    workload description per mode; no GPU needed."""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    spec.loader.exec_module(b)
    a = b.parse()
    assert (a.gpus, a.impl, a.mode) == (1, "b200", "train") and a.warmup >= 3 and a.steps > 0
    cfg = b.workload_config(a, a.mode, a.classes, a.distill, a.images_per_gpu)
    assert "configs[1]" in cfg["workload"] and cfg["proposals_per_image"] == 512 and "model" not in cfg
    assert "forward(" in cfg["workload"] and cfg["api"] == "forward()" and cfg["rpn_proposals_per_image"] == 2000
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "8", "--steps", "7", "--warmup", "4", "--impl", "reference"])
    a = b.parse()
    assert (a.gpus, a.steps, a.warmup, a.impl) == (8, 7, 4, "reference")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--distill"])
    a = b.parse()
    assert "configs[3]" in b.workload_config(a, a.mode, a.classes, a.distill, 8)["workload"]
    monkeypatch.setattr(sys, "argv", ["bench.py", "--mode", "infer", "--classes", "80"])
    a = b.parse()
    assert "inference step" in b.workload_config(a, a.mode, a.classes, a.distill, 8)["workload"] and a.classes == 80
    assert b.METRIC == "roi_head_images_per_sec" and b.UNIT == "images/s"


def test_pcb_constructs_like_the_reference_and_preprocesses_like_it():
    """`PrototypicalCalibrationBlock(cfg)` — the reference's one-argument call (evaluator.py:90): ImageNet ResNet-101 with
    torchvision's parameter names (so cfg.TEST.PCB_MODELPATH loads), the reference's preprocessing (calibration_layer.py:
    91-98) and the reference's dataset-dict support format; no GPU needed up to the pooling."""
    import torchvision
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    from fewshotobjectdetection_imporove_via_text_feature_b200.evaluation.calibration_layer import PrototypicalCalibrationBlock
    cfg = config.get_cfg()
    cfg.MODEL.DEVICE = "cpu"
    torch.manual_seed(0)
    pcb = PrototypicalCalibrationBlock(cfg)
    ref = torchvision.models.resnet101()
    assert list(pcb.imagenet_model.state_dict().keys()) == list(ref.state_dict().keys())
    assert all(a.shape == b.shape for a, b in zip(pcb.imagenet_model.state_dict().values(), ref.state_dict().values()))
    assert pcb.fc is pcb.imagenet_model.fc and pcb.exclude_cls == list(range(15))      # voc test_all: base classes excluded
    # same network as torchvision's, fed the reference's preprocessing
    ref.load_state_dict(pcb.imagenet_model.state_dict())
    ref.eval()
    img = np.random.RandomState(0).randint(0, 256, (96, 128, 3)).astype(np.uint8)       # BGR, as cv2.imread returns it
    with torch.no_grad():
        got = pcb.feature_extractor(img)
        mean = torch.tensor([0.406, 0.456, 0.485]).reshape(3, 1, 1)
        std = torch.tensor([0.225, 0.224, 0.229]).reshape(3, 1, 1)
        x = ((torch.from_numpy(img.transpose(2, 0, 1)) / 255.0 - mean) / std)[None][:, [2, 1, 0]]
        r = ref
        want = r.layer4(r.layer3(r.layer2(r.layer1(r.maxpool(r.relu(r.bn1(r.conv1(x))))))))
    assert got.shape == (1, 2048, 3, 4)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_gemm2_descriptor_fields_agree_everywhere():
    """struct b200_gemm2_desc: the header's field list == the package's ctypes mirror == the binding INTEGRATION.md shows a
    maintainer (names, order and size; a silent mismatch would shift every later field)."""
    import ctypes
    import re
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200roi.h")).read()
    body = hdr[hdr.index("typedef struct b200_gemm2_desc {"):hdr.index("} b200_gemm2_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if stmt:
            parts = stmt.split(",")
            names.append(parts[0].split()[-1].lstrip("*"))
            names += [q.strip().lstrip("*") for q in parts[1:]]
    mirror = [n for n, _ in _lib.Gemm2Desc._fields_]
    assert names == mirror, (names, mirror)
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    block = next(b for b in re.findall(r"```python\n(.*?)```", md, re.S) if "ctypes.CDLL" in b)
    ns = {"ctypes": ctypes}
    exec(block[block.index("class Gemm2Desc"):block.index("def linear_relu")], ns)
    assert [n for n, _ in ns["Gemm2Desc"]._fields_] == mirror
    assert ctypes.sizeof(ns["Gemm2Desc"]) == ctypes.sizeof(_lib.Gemm2Desc)


def test_cat_adjacent_is_a_view_only_when_the_pieces_sit_back_to_back():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    base = torch.arange(24.0).reshape(6, 4)
    parts = [base[0:2], base[2:5], base[5:6]]
    v = ops.cat_adjacent(parts)
    assert torch.equal(v, base) and v.data_ptr() == base.data_ptr()                     # view: no copy
    gap = [base[0:2], base[3:5]]
    c = ops.cat_adjacent(gap)
    assert torch.equal(c, torch.cat(gap)) and c.data_ptr() != base.data_ptr()            # hole -> real concatenation
    other = [base[0:2], torch.ones(3, 4)]
    assert torch.equal(ops.cat_adjacent(other), torch.cat(other))
    assert torch.equal(ops.cat_adjacent([base[1:3]]), base[1:3])


def test_split_losses_backward_equals_indexing():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import train_ops
    w = torch.tensor([0.5, 2.0, -1.0, 3.0])
    x1 = torch.tensor([1.0, 2.0, 3.0, 4.0], requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    a = train_ops.split_losses(x1 * w)
    (a[0] + 3 * a[2] + a[3]).backward()                     # a[1] unused: its gradient arrives as None
    y = x2 * w
    (y[0] + 3 * y[2] + y[3]).backward()
    assert torch.equal(x1.grad, x2.grad)


def test_text_domination_operand_layout():
    """300-d operands of LV_attention_textDomination in pitch-304 / 608 buffers: the padded weights reproduce linear3 / proj_value
    on the padded concatenations exactly (zero weight columns under the gaps)."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, ops
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling import roi_heads as RH
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    torch.manual_seed(0)
    m = RH.LV_attention_textDomination(64, cfg=cfg, class_embed=torch.randn(5, 300)).eval()
    named = {k: v for k, v in m.named_parameters() if not k.startswith(("mlp_adapter", "proj_k"))}
    w = ops.TextDominationWeights().refresh(named, (m.embed,))
    d, h, dp, hp = w["d"], w["h"], w["dp"], w["hp"]
    assert (d, h, dp, hp) == (300, 150, 304, 152)
    o1, o2, q = torch.randn(7, h), torch.randn(7, h), torch.randn(7, d)
    xcat = torch.zeros(7, 2 * dp)
    xcat[:, :h], xcat[:, hp:hp + h], xcat[:, 2 * hp:2 * hp + d] = o1, o2, q
    assert w["linear3.weight"].shape == (d, 2 * hp + d) and w["linear3.weight"].stride(0) == 2 * dp
    W3 = w["linear3.weight"].float()
    ref = torch.cat([o1, o2, q], 1) @ m.attention.linear3.weight.detach().to(torch.bfloat16).float().t()
    torch.testing.assert_close(xcat[:, :2 * hp + d] @ W3.t(), ref, rtol=1e-5, atol=1e-5)
    v300, t = torch.randn(7, d), torch.randn(7, d)
    cat = torch.zeros(7, 2 * dp)
    cat[:, :d], cat[:, dp:dp + d] = v300, t
    ref = torch.cat([v300, t], 1) @ m.proj_value.weight.detach().to(torch.bfloat16).float().t()
    torch.testing.assert_close(cat[:, :dp + d] @ w["proj_value.weight"].float().t(), ref, rtol=1e-5, atol=1e-5)
    for k in ("linear1.weight", "linear2.weight", "ffn1.weight", "proj2.weight", "kq"):
        assert w[k].stride(0) % 8 == 0 and w[k].stride(1) == 1, k


def test_roi_bwd_impl_default_mirror():
    """_lib.ROI_BWD_IMPL_DEFAULT (what tests restore "roi_align_bwd_impl" to) == the initial value in the CUDA source."""
    import re
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib
    src = open(os.path.join(os.path.dirname(_lib.__file__), "csrc", "roi_align_bwd.cu")).read()
    m = re.search(r"^int g_roi_bwd_impl = (\d+);", src, re.M)
    assert m and int(m.group(1)) == _lib.ROI_BWD_IMPL_DEFAULT
