#!/usr/bin/env python
"""bench.py — ROI-head images/s on B200 (BASELINE.json metric), with roofline, CPU baseline and e2e figures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode train|infer]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload = BASELINE.json configs[1]: "VOC split1 10-shot novel fine-tune with CLIP text-fused ROI head, bf16,
1 B200".  A "step" is one fine-tune step of the text-fused C4 ROI head over one batch of synthetic input on each GPU,
driven through the reference-facing API exactly as defrcn/modeling/meta_arch/rcnn.py:94-98 drives it:

    f = affine_rcnn(decouple_layer(features["res4"], lambda))            # G1 + G2 (caller side of the head)
    _, losses = roi_heads(images, {"res4": f}, proposals, targets)        # SematicRes5ROIHeads.forward(...)
    sum(losses.values()).backward();  [gradient all-reduce];  SGD + momentum

with 8 images x 2000 RPN-like proposals + 8 ground-truth boxes per image (600x800 px -> res4 38x50x1024, K = 20, CLIP
512-d).  Inside `forward`: label_and_sample_proposals (S1: 2008 candidates -> 512 sampled rows per image) -> ROIAlign ->
res5 (frozen, own tcgen05 kernels) -> mean -> text fusion -> cls_score (dropout 0.8) / bbox_pred -> the three losses.
`--mode infer` times the eval-mode forward (512 proposals per image -> softmax/decode/threshold/per-class NMS/top-100).
Images shard across GPUs (weak scaling); the only exchanges are the gradient all-reduce (train) / the detection
all-gather (infer).  After the headline measurement the same process also measures BASELINE configs[2] / [3] and the
inference direction briefly and attaches them as `extra` (skip with --no-extras).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IMG, W_IMG, HF, WF, C4 = 600, 800, 38, 50, 1024
METRIC, UNIT = "roi_head_images_per_sec", "images/s"
GDL_LAMBDA, DROP_P, LR, MOMENTUM, WD = 0.001, 0.8, 0.01, 0.9, 5e-5   # configs/voc/defrcn_fsod_r101_novel*: fine-tune


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--images-per-gpu", type=int, default=8)
    ap.add_argument("--props", type=int, default=512, help="ROIs per image through the head (sampled rows in training)")
    ap.add_argument("--rpn-props", type=int, default=2000, help="training: unsampled proposals per image handed to forward()")
    ap.add_argument("--gt-per-image", type=int, default=8)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--cpu-baseline-images", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short configs[2] / [3] / inference passes")
    ap.add_argument("--distill", action="store_true",
                    help="train mode: BASELINE configs[3] — the student step with the KL loss against a frozen VKV teacher "
                         "(SematicRes5ROIHeadsDistill); not the default metric line")
    ap.add_argument("--no-graph", action="store_true",
                    help="fine-tune mode: enqueue every step from the host instead of replaying one captured CUDA graph")
    return ap.parse_args()


def synth_inputs(n_images, props, seed0=1234, num_classes=20, n_obj=8):
    """SURVEY.md §8(d) synthetic inputs: post-ReLU res4 maps; per image `props` RPN-like + jittered proposals around
    `n_obj` objects, which are also the image's ground truth (random classes); per-image seeds."""
    from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals
    feat = torch.relu(torch.randn(n_images, C4, HF, WF, generator=torch.Generator().manual_seed(0)))
    boxes, gt_cls, gt_boxes = [], [], []
    for i in range(n_images):
        gen = torch.Generator().manual_seed(seed0 + i)
        b, objs = synth_proposals(props, H_IMG, W_IMG, gen, n_obj=n_obj)
        objs = objs.clone()
        objs[:, 2:] = torch.maximum(objs[:, 2:], objs[:, :2] + 2)
        boxes.append(b)
        gt_boxes.append(objs)
        gt_cls.append(torch.randint(0, num_classes, (n_obj,), generator=gen))
    return feat, boxes, gt_cls, gt_boxes


def build_head(num_classes, device, train=False, distill=False, props=512):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = "SematicRes5ROIHeadsDistill" if (train and distill) else "SematicRes5ROIHeads"
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = num_classes
    cfg.MODEL.ADDITION.NAME = "clip"
    if train:
        cfg.MODEL.ROI_HEADS.CLS_DROPOUT = True
        cfg.MODEL.ROI_HEADS.DROPOUT_RATIO = DROP_P
        cfg.MODEL.ROI_HEADS.ENABLE_DECOUPLE = True
        cfg.MODEL.ROI_HEADS.BACKWARD_SCALE = GDL_LAMBDA
        cfg.MODEL.ROI_HEADS.FREEZE_FEAT = True
        cfg.MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE = props
        cfg.MODEL.B200.STATIC_SAMPLING = True      # no host read inside forward(): the step is captured in one CUDA graph
    torch.manual_seed(0)
    head = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=C4, stride=16)})
    head = head.train() if train else head.eval()
    aff = modeling.AffineLayer(C4, bias=True)
    with torch.no_grad():
        # random-init weights of the reference architecture; classifier scaled so that scores are not uniform
        head.box_predictor.cls_score.weight.mul_(40.0)
        head.box_predictor.bbox_pred.weight.mul_(50.0)
        if train and distill:
            head.teacher_cls_score.weight.mul_(40.0)
        aff.weight.normal_(1.0, 0.05)
        aff.bias.normal_(0.0, 0.05)
    if train:
        for p in head.res5.parameters():            # ROI_HEADS.FREEZE_FEAT (roi_heads.py:96-99)
            p.requires_grad = False
    return cfg, head.to(device), aff.to(device)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference modules' arithmetic + torchvision CPU ops) on this box's host cores
# ---------------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, args):
        torch.set_num_threads(os.cpu_count() or 1)
        self.args, self.train = args, args.mode == "train"
        _, head, aff = build_head(args.classes, "cpu", train=self.train, props=args.props)
        self.params = {k: v.detach().float().clone() for k, v in head.state_dict().items()}
        self.text = torch.cat([head.attention.embed, head.attention.bg_feature], 0).float()
        self.aff_w, self.aff_b = aff.weight.detach().clone(), aff.bias.detach().clone()
        n_in = args.rpn_props if self.train else args.props
        self.feat, self.boxes, self.gt_cls, self.gt_boxes = synth_inputs(1, n_in, num_classes=args.classes, n_obj=args.gt_per_image)
        self.gen = torch.Generator().manual_seed(5)
        if self.train:
            self.trainable = [v for k, v in self.params.items() if k.startswith(("attention.", "box_predictor."))]
            self.trainable += [self.aff_w, self.aff_b]
            for t in self.trainable:
                t.requires_grad_(True)
            self.opt = torch.optim.SGD(self.trainable, lr=LR, momentum=MOMENTUM, weight_decay=WD)

    def step(self, stages=None):
        from oracle import oracle as O
        a = self.args
        if not self.train:
            with torch.no_grad():
                return O.head_forward(self.feat * self.aff_w + self.aff_b, self.boxes, [(H_IMG, W_IMG)], self.text, self.params,
                                      stages=stages)
        self.opt.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        with torch.no_grad():        # S1: roi_heads.py:157-250
            b, c, g = O.label_and_sample(self.boxes[0], self.gt_boxes[0], self.gt_cls[0], a.classes, batch=a.props, gen=self.gen)
        if stages is not None:
            stages["label_sample"] = stages.get("label_sample", 0.0) + time.perf_counter() - t0
        out = O.head_train_step(self.feat, [b], c, g, self.text, self.params, self.aff_w, self.aff_b, GDL_LAMBDA, a.classes,
                                DROP_P, stages=stages)
        self.opt.step()
        return out

    def describe(self, n):
        return "%d image(s) x %d proposals%s, %s step, fp32, torch %d threads" % (
            n, self.args.props, " sampled from %d" % self.args.rpn_props if self.train else "",
            "fine-tune (label+sample, fwd, bwd, SGD)" if self.train else "inference", torch.get_num_threads())


def run_cpu_baseline(args, warm=1):
    arm = CpuArm(args)
    for _ in range(warm):
        arm.step()
    stages, t0 = {}, time.perf_counter()
    for _ in range(args.cpu_baseline_images):
        arm.step(stages)
    dt = time.perf_counter() - t0
    return {"value": args.cpu_baseline_images / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": arm.describe(args.cpu_baseline_images) + ", after %d warm-up" % warm,
            "stage_ms_per_image": {k: 1e3 * v / args.cpu_baseline_images for k, v in stages.items()}}


def main_reference(args, rank, world):
    if rank != 0:
        return
    arm = CpuArm(args)
    for _ in range(args.warmup):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cb = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
          "sample": "each step = " + arm.describe(1) + " (bounded sample of the same workload)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the own arm's workload; each timed step of this arm is the bounded sample `cpu_baseline.sample` names
        "config": workload_config(args, args.mode, args.classes, args.distill, args.images_per_gpu), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, mode, classes, distill, images_per_gpu):
    if mode == "train":
        what = ("fine-tune step through the public API (BASELINE configs[%d]): affine_rcnn(decouple_layer(res4)) -> "
                "SematicRes5ROIHeads%s.forward(images, features, proposals, targets) [label_and_sample_proposals %d+%d -> %d rows/image, "
                "ROIAlign 7x7, res5 (frozen), text fusion, cls_score(dropout 0.8)/bbox_pred, loss_cls+loss_box_reg+loss_attentive%s] "
                "-> backward to the res4 map and all trained parameters -> (grad all-reduce) -> SGD+momentum"
                % (3 if distill else 1, "Distill" if distill else "", args.rpn_props, args.gt_per_image, args.props,
                   "+loss_kl vs the frozen LV_attention_VKV teacher" if distill else ""))
    else:
        what = ("inference step through the public API: affine_rcnn(res4) -> SematicRes5ROIHeads.forward(images, features, proposals) "
                "[ROIAlign 7x7, res5, text fusion, softmax/decode/threshold/per-class NMS/top-100] -> Instances")
    return {"workload": "DeFRCN R-101 C4 text-fused ROI head (CLIP 512-d, K=%d), %s" % (classes, what),
            "mode": mode, "classes": classes, "api": "forward()",
            "roi_align_bins": "16 live of 49 (stride-2 consumer)",
            "images_per_gpu_per_step": images_per_gpu, "proposals_per_image": args.props,
            "rpn_proposals_per_image": args.rpn_props if mode == "train" else None, "image_px": [H_IMG, W_IMG],
            "res4_map": [C4, HF, WF], "l2": "flushed between timed steps (256 MiB write)", "parallelism": "image-sharded dp%d" % args.gpus}


TRAIN_STAGES = ["gdl_affine", "label_sample", "roi_align", "res5_mean", "text_fusion_losses", "bwd_text_fusion", "bwd_res5",
                "bwd_roi_align", "bwd_gdl_affine", "allreduce_sgd"]
INFER_STAGES = ["affine", "roi_align", "res5_mean", "text_fusion_predictor", "decode_nms"]


class Workload:
    """One configuration of the head on this rank's GPU: synthetic inputs (pinned host + device-resident copies), the
    step function through `forward()`, CUDA-graph capture of the fine-tune step, device-timed and end-to-end loops."""

    def __init__(self, args, mode, classes, distill, dev, rank, world, graph=True):
        from fewshotobjectdetection_imporove_via_text_feature_b200 import train_ops
        self.args, self.mode, self.K, self.distill, self.dev, self.rank, self.world = args, mode, classes, distill, dev, rank, world
        self.train = mode == "train"
        self.B, self.P = args.images_per_gpu, args.props
        self.cfg, self.head, self.aff = build_head(classes, dev, train=self.train, distill=distill, props=self.P)
        n_in = args.rpn_props if self.train else self.P
        feat_h, boxes_h, cls_h, gtb_h = synth_inputs(self.B, n_in, seed0=1234 + 1000 * rank, num_classes=classes, n_obj=args.gt_per_image)
        self.host = {"feat": feat_h.pin_memory(), "boxes": torch.stack(boxes_h).pin_memory()}
        if self.train:
            self.host["gt_cls"] = torch.stack(cls_h).pin_memory()
            self.host["gt_boxes"] = torch.stack(gtb_h).pin_memory()
        self.host["objectness"] = torch.zeros(self.B, n_in).pin_memory()          # RPN objectness logits travel with the boxes
        self.names = list(self.host)
        self.resident = {k: v.to(dev) for k, v in self.host.items()}
        self.sizes = [(H_IMG, W_IMG)] * self.B
        self.opt = None
        if self.train:
            self.opt = train_ops.FlatSGD(list(self.head.attention.parameters()) + list(self.head.box_predictor.parameters()) +
                                         list(self.aff.parameters()), lr=LR, momentum=MOMENTUM, weight_decay=WD, direct_grads=True)
        self.stage_names = TRAIN_STAGES if self.train else INFER_STAGES
        self.n_marks = len(self.stage_names) + 1
        self.graph, self.graph_note, self.in_graph = None, "off (--no-graph)" if self.train else "n/a (forward() returns Instances: one host read)", False
        self.want_graph = graph and self.train and not args.no_graph

    # -- inputs in the reference's containers ---------------------------------------------------------------------
    def batch(self, d):
        from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances
        props, targets = [], ([] if self.train else None)
        for i in range(self.B):
            inst = Instances(self.sizes[i])
            inst.proposal_boxes = Boxes(d["boxes"][i])
            inst.objectness_logits = d["objectness"][i]
            props.append(inst)
            if self.train:
                t = Instances(self.sizes[i])
                t.gt_boxes = Boxes(d["gt_boxes"][i])
                t.gt_classes = d["gt_cls"][i]
                targets.append(t)
        return props, targets

    # -- one step through the public API --------------------------------------------------------------------------
    def step(self, d, ev=None, exchange=True):
        import torch.distributed as dist
        head, aff, opt = self.head, self.aff, self.opt
        idx = {n: i + 1 for i, n in enumerate(self.stage_names)}

        def mark(i):
            if ev is not None:
                ev[i].record()

        def cb(name, tensor):
            if name in idx:
                mark(idx[name])
            if self.train and ev is not None and tensor is not None and tensor.requires_grad:
                if name == "res5_mean":
                    tensor.register_hook(lambda g: ev[idx["bwd_text_fusion"]].record())
                elif name == "roi_align":
                    tensor.register_hook(lambda g: ev[idx["bwd_res5"]].record())
        head._stage_cb = cb if ev is not None else None
        props, targets = self.batch(d)
        mark(0)
        if not self.train:
            f = aff(d["feat"], None, True, torch.bfloat16)                                # G2 (+ layout / dtype for the gather)
            mark(1)
            pred, _ = head(None, {"res4": f}, props, None)                                # forward(): P1, P2, T/A/C, D1-D3
            return pred
        opt.zero_grad()
        x = d["feat"].detach().requires_grad_(True)
        f = aff(x, GDL_LAMBDA, True, torch.bfloat16)                                      # G1 + G2 (rcnn.py:94-97)
        mark(1)
        if ev is not None:
            f.register_hook(lambda g: ev[idx["bwd_roi_align"]].record())
        _, losses = head(None, {"res4": f}, props, targets)                               # forward(): S1, P1, P2, T/A/C, L1
        total = losses["loss_cls"] + losses["loss_box_reg"] + losses["loss_attentive"]
        if self.distill:
            total = total + losses["loss_kl"]
        total.backward()                                                                  # L1, A*, P2, P1b, G1/G2 backward
        mark(idx["bwd_gdl_affine"])
        res = {"losses": torch.stack([losses["loss_cls"], losses["loss_box_reg"], losses["loss_attentive"]]).detach(),
               "grad_feat": x.grad}
        if not exchange:
            opt.sync_grads()         # join the parameter-gradient streams (a captured region must end on one stream)
            return res
        if self.world > 1:
            # the one real exchange step: head gradients on a communication stream behind the parameter-gradient streams
            # (under the res5 / ROIAlign backward still queued on the GPU), affine_rcnn's two vectors at the end
            opt.all_reduce_grads(n_late_params=len(list(aff.parameters())))
        opt.step()
        mark(idx["allreduce_sgd"])
        return res

    def prepare(self, warm):
        """Warm-up steps, then (fine-tune) capture of the whole step into one CUDA graph."""
        from fewshotobjectdetection_imporove_via_text_feature_b200 import train_ops
        with (torch.enable_grad() if self.train else torch.no_grad()):
            for _ in range(max(warm, 3)):
                self.out = self.step(self.resident)
            torch.cuda.synchronize()
            if self.train:
                self.head.flush_deferred_logs()          # raises if static sampling did not hold on this batch
            if self.want_graph:
                try:
                    self.head.use_device_dropout_counter(True)
                    # world > 1: the NCCL gradient all-reduce is captured too (on the communication stream, behind the
                    # parameter-gradient streams and under res5's backward); BENCH_GRAPH_ALLREDUCE=0 keeps it and SGD outside
                    self.in_graph = self.world == 1 or os.environ.get("BENCH_GRAPH_ALLREDUCE", "1") == "1"
                    self.graph = train_ops.GraphedStep(lambda d: self.step(d, exchange=self.in_graph), self.resident)
                    self.graph_note = "whole step" if self.in_graph else "forward + backward (all-reduce and SGD outside)"
                    # the resident inputs of the device-timed loop ARE the graph's static buffers: no per-step device copy
                    self.resident = dict(self.graph.static_in)
                except Exception as e:  # noqa: BLE001
                    self.graph, self.graph_note = None, "capture failed, running eagerly: %s" % str(e).splitlines()[0][:200]
                    self.head.use_device_dropout_counter(False)
                    torch.cuda.synchronize()

    def run_step(self, d, ev=None):
        import torch.distributed as dist
        if self.graph is None:
            with (torch.enable_grad() if self.train else torch.no_grad()):
                return self.step(d, ev)
        if ev is not None:
            ev[0].record()
        static_out = self.graph(d)
        if self.world > 1 and not self.in_graph:
            dist.all_reduce(self.opt.grad, op=dist.ReduceOp.AVG)
            self.opt.step()
        if ev is not None:
            ev[self.n_marks - 1].record()
        return static_out

    def timed(self, steps, flush):
        """Device-timed loop: `steps` steps, L2 flushed before each, one event pair per step (first mark .. last mark);
        returns (total ms max over ranks, per-step list, kernel launches per step, host enqueue ms per step)."""
        import torch.distributed as dist
        from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(self.n_marks)] for _ in range(steps)]
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.LAUNCHES
        t0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xff)
            self.out = self.run_step(self.resident, evs[i])
            if not self.train:
                evs[i][self.n_marks - 1].record()
        host_ms = (time.perf_counter() - t0) * 1e3 / max(steps, 1)
        torch.cuda.synchronize()
        launches = (_lib.LAUNCHES - l0) // max(steps, 1)
        if self.graph is not None:       # replays bypass the Python entry points: count the kernels recorded into the graph
            launches = self.graph.kernel_launches
        if self.world > 1:
            dist.barrier()
        per_step = [evs[i][0].elapsed_time(evs[i][self.n_marks - 1]) for i in range(steps)]
        total = torch.tensor([float(sum(per_step))], device=self.dev)
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        if self.train:
            self.head.flush_deferred_logs()
        return float(total), per_step, int(launches), host_ms, evs

    def stage_split(self, flush, evs=None, n=5):
        """Per-stage times from eager steps with marks at the stage boundaries inside forward() / the backward."""
        if evs is None or self.graph is not None:
            with (torch.enable_grad() if self.train else torch.no_grad()):
                for _ in range(4):           # the eager path's allocations settle again after the capture
                    self.step(self.resident)
                evs = [[torch.cuda.Event(enable_timing=True) for _ in range(self.n_marks)] for _ in range(n)]
                for i in range(n):
                    flush.fill_(i)
                    self.step(self.resident, evs[i])
                    if not self.train:
                        evs[i][self.n_marks - 1].record()
            torch.cuda.synchronize()
            self.head._stage_cb = None
        return dict(zip(self.stage_names, [float(np.median([e[s].elapsed_time(e[s + 1]) for e in evs])) for s in range(self.n_marks - 1)]))

    # -- end to end: pinned host inputs -> device, result -> host, every step --------------------------------------
    def result_tensors(self, out):
        if self.train:
            return {"losses": out["losses"]}
        return {"boxes": torch.cat([r.pred_boxes.tensor for r in out], 0), "scores": torch.cat([r.scores for r in out], 0),
                "classes": torch.cat([r.pred_classes for r in out], 0)}

    def e2e(self, steps):
        import torch.distributed as dist
        dev, host, names = self.dev, self.host, self.names
        cap = {"losses": (3,), "boxes": (self.B * 100, 4), "scores": (self.B * 100,), "classes": (self.B * 100,)}
        ex = self.result_tensors(self.out)
        res_host = {k: torch.empty(cap[k], dtype=v.dtype).pin_memory() for k, v in ex.items()}
        cur = torch.cuda.current_stream()
        cpy = torch.cuda.Stream()
        dbuf = [{k: torch.empty_like(self.resident[k]) for k in names} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        d2h = [0]

        def upload(i):
            b = i & 1
            with torch.cuda.stream(cpy):
                cpy.wait_event(free[b])
                for k in names:
                    dbuf[b][k].copy_(host[k], non_blocking=True)
                ready[b].record(cpy)

        def run(n, overlap):
            for e in free:
                e.record(cur)
            if overlap:
                upload(0)
            for i in range(n):
                b = i & 1
                if overlap:
                    if i + 1 < n:
                        upload(i + 1)
                    cur.wait_event(ready[b])
                    o = self.run_step(dbuf[b])
                    free[b].record(cur)
                else:
                    o = self.run_step({k: host[k].to(dev, non_blocking=True) for k in names})
                r = self.result_tensors(o)
                d2h[0] = 0
                for k, v in r.items():
                    res_host[k][: v.shape[0]].copy_(v, non_blocking=True)
                    d2h[0] += v.numel() * v.element_size()
                if not overlap:
                    cur.synchronize()
            cur.synchronize()

        out = {}
        for name, overlap in (("serial", False), ("overlap", True)):
            run(3, overlap)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            run(steps, overlap)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if self.world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out[name] = float(ms)
        return out, int(sum(v.numel() * v.element_size() for v in host.values())), int(d2h[0])

    def close(self):
        self.graph = None
        self.head._stage_cb = None


def replay_gemms(cases, dev, flush):
    """Every tensor-core GEMM launch of one step (same shapes, operand layouts and epilogues; fresh operands) re-issued
    back to back on the launching stream inside ONE CUDA event pair, L2 flushed before the sequence: the sum of the
    kernels' durations without the cross-stream SM sharing of the step and without per-launch host latency."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops, train_ops
    calls, flop = [], 0.0
    rnd0 = lambda *s, sc=0.5: (torch.randn(*s, device=dev) * sc).to(torch.bfloat16)
    rnd = lambda r, c_, sc=0.5: rnd0(r, (c_ + 7) // 8 * 8, sc=sc)[:, :c_]          # row pitch a multiple of 8 elements (TMA)
    for c in cases:
        if isinstance(c, dict):                      # b200_gemm2
            M, N, K, K2 = c["M"], c["N"], c["K"], c["K2"]
            if c["conv_c"]:
                a = rnd(M, c["conv_c"])
            else:
                a = rnd(K, M) if c["a_mn"] else rnd(M, K)
            b = rnd(K, N, sc=0.05) if c["b_mn"] else rnd(N, K + K2, sc=0.05)
            kw = dict(a_mn=c["a_mn"], b_mn=c["b_mn"], conv_c=c["conv_c"], relu=c["relu"], accumulate=c["acc"], want_out=c["out"])
            if K2:
                kw["a2"] = rnd(M, K2)
            if c["bias"]:
                kw["bias"] = torch.randn(N, device=dev)
            if c["res"]:
                kw["residual"] = rnd(M, N)
            if c["mbits"]:
                kw["mask_bits"] = torch.randint(-2 ** 31, 2 ** 31 - 1, (M, (N + 31) // 32), dtype=torch.int32, device=dev)
            if c["mact"]:
                kw["mask_act"] = rnd(M, N)
            if c["out"]:
                kw["out"] = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device=dev)[:, :N]
            if c["out2"]:
                kw["out2"] = torch.empty(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device=dev)[:, :N]
            if c["f32"]:
                kw["out_f32"] = torch.zeros(M, N, device=dev)
            if c["bout"]:
                kw["bits_out"] = torch.empty(M, N // 32, dtype=torch.int32, device=dev)
            if c["mean"]:
                kw["rowmean_out"] = torch.empty(M // 16, N, device=dev)
            if c.get("ssq"):
                kw["rowsumsq_out"] = torch.empty(M, (N + 63) // 64, device=dev)
            if c.get("rscale"):
                kw["row_scale_sumsq"] = torch.rand(M, c["rscale"], device=dev) + 0.5
            if c.get("softmax"):
                kw["softmax"] = True
            if c.get("gate"):
                kw["gate"] = True
            calls.append((ops.gemm2, (a, b), kw))
            flop += 2.0 * M * N * (K + K2)
        else:                                        # b200_gemm_bf16(_ex), single-CTA kernel
            (M, N, K, obf, d2, relu, acc, msk, bias) = c
            ld = (N + 7) // 8 * 8
            kw = dict(relu=relu, accumulate=acc,
                      out=torch.zeros(M, ld, device=dev, dtype=torch.bfloat16 if obf else torch.float32)[:, :N],
                      out2=torch.empty(M, ld, device=dev, dtype=torch.bfloat16)[:, :N] if d2 else None,
                      mask=rnd(M, ld)[:, :N] if msk else None)
            calls.append((train_ops.gemm_ex, (rnd(M, K), rnd(N, K, sc=0.05), torch.randn(N, device=dev) if bias else None), kw))
            flop += 2.0 * M * N * K
    ts = []
    for it in range(6):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for fn, a, kw in calls:
            fn(*a, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:])) if calls else float("nan"), flop, len(calls)


def traffic_from_profiles(key):
    """DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` summary of
    this exact configuration (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic.py); None when not captured."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"))).get(key)
    except Exception:  # noqa: BLE001
        return None


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank, world)

    import torch.distributed as dist
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, distributed as bdist, ops

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NB the GPU box exports NCCL_DEBUG=VERSION: NCCL itself prints one "NCCL version ..." line to stdout before the JSON line
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, P, K = args.images_per_gpu, args.props, args.classes
    train = args.mode == "train"

    wl = Workload(args, args.mode, K, train and args.distill, dev, rank, world)
    wl.prepare(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    total_ms, per_step, launches, host_ms, evs = wl.timed(args.steps, flush)
    sampler.stop_flag = True
    stage_ms = wl.stage_split(flush, evs)

    # ---- per-entry-point profile (separate eager pass: an event pair around every C-ABI call) ---------------------
    _lib.PROFILE = {}
    prof_steps = 5
    with (torch.enable_grad() if train else torch.no_grad()):
        for i in range(prof_steps):
            flush.fill_(i)
            wl.step(wl.resident)
    torch.cuda.synchronize()
    prof, gemm_cases = {}, []
    fl_of = lambda t: (t[0] if isinstance(t, tuple) else t) or 0.0
    for name, rows in _lib.PROFILE.items():
        per = max(len(rows) // prof_steps, 1)            # the same calls every step: median over the profiled steps
        ms = float(np.median([sum(a.elapsed_time(b) for a, b, _ in rows[i * per:(i + 1) * per]) for i in range(prof_steps)]))
        fl = sum(fl_of(t) for _, _, t in rows if t) / prof_steps
        prof[name] = {"ms_per_step": ms, "calls_per_step": len(rows) / prof_steps}
        if fl:
            prof[name]["tflops"] = fl / (ms * 1e-3) / 1e12
            prof[name]["flop_per_step"] = fl
        if name in ("b200_gemm_bf16", "b200_gemm_bf16_ex", "b200_gemm2"):
            gemm_cases += [t[1] for _, _, t in rows[:per] if isinstance(t, tuple)]
    if os.environ.get("BENCH_DUMP_CALLS"):
        for name, rows in _lib.PROFILE.items():
            for a, b, t in rows[: len(rows) // prof_steps]:
                ms = a.elapsed_time(b)
                print("CALL %-34s %8.4f ms %s" % (name, ms, ("%.1f GF  %.0f TF/s %s" % (fl_of(t) / 1e9, fl_of(t) / ms / 1e9, t[1] if isinstance(t, tuple) else "")) if t else ""), file=sys.stderr)
    _lib.PROFILE = None
    b2b_ms, gemm_flop, gemm_calls = replay_gemms(gemm_cases, dev, flush)

    # ---- the stand-alone ROIAlign operator (all 49 bins, what torchvision.ops.roi_align computes) on the same maps, and
    # its backward (lists planned ahead, as in the step), L2 flushed before every launch
    with torch.enable_grad():
        fmap = wl.resident["feat"].to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        rois_l = [wl.resident["boxes"][i][:P] for i in range(B)]
        rr, oo = ops.boxes_to_rois(rois_l)
        tf, tb, ts16 = [], [], []
        for i in range(5):
            flush.fill_(i)
            e0_, e1_, e2_, e3_ = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0_.record()
            pooled_full = ops.roi_align(fmap, rr, 7, 1.0 / 16, 0, True, channels_last_out=True, roi_batch_offsets=oo, bin_step=1)
            e1_.record()
            gfull = torch.ones_like(pooled_full)
            torch.cuda.synchronize()      # the plan (side stream, forked by the forward) is complete: "planned ahead", as in
            flush.fill_(i + 1)            # the step, where it runs under res5 — the timed region is the gather launch alone
            e2_.record()
            pooled_full.backward(gfull)
            e3_.record()
            flush.fill_(i + 2)
            e4_, e5_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e4_.record()
            with torch.no_grad():
                ops.roi_align(fmap, rr, 7, 1.0 / 16, 0, True, channels_last_out=True, roi_batch_offsets=oo, bin_step=2)
            e5_.record()
            torch.cuda.synchronize()
            fmap.grad = None
            if i >= 2:
                tf.append(e0_.elapsed_time(e1_))
                tb.append(e2_.elapsed_time(e3_))
                ts16.append(e4_.elapsed_time(e5_))
        full_bytes = B * C4 * HF * WF * 2 + B * P * 20 + B * P * C4 * 49 * 2
        live_bytes = B * C4 * HF * WF * 2 + B * P * 20 + B * P * C4 * 16 * 2
        roi_op = {"bins": "7x7", "algorithmic_bytes": full_bytes,
                  "fwd_ms": float(np.median(tf)), "fwd_gbs": full_bytes / float(np.median(tf)) / 1e6,
                  "bwd_ms": float(np.median(tb)), "bwd_gbs": full_bytes / float(np.median(tb)) / 1e6,
                  "bwd_kernel": "roi_bwd_tile_gather_kernel<1,false,3,8> (4x4-pixel tiles, mma.sync m16n8k16 bf16; plan = "
                                "roi_slice_prepare_kernel + roi_bwd_tile_build_kernel, built ahead on the plan stream and complete "
                                "before the timed launch)",
                  "note": "b200_roi_align_fwd / b200_roi_align_bwd_planned entry points, CUDA events, median of 3"}
        roi16_ms = float(np.median(ts16))
        del pooled_full, gfull, fmap
    # ---- the post-processing operator (BASELINE metric "NMS us"): softmax + decode + threshold compaction, per-class NMS,
    # top-100 gather on this batch's proposals with SURVEY 8(d)'s logits, L2 flushed before every call
    with torch.no_grad():
        gen_ = torch.Generator().manual_seed(99)
        lg_ = torch.randn(B * P, K + 1, generator=gen_)
        peak_ = torch.rand(B * P, generator=gen_) < 0.3
        cls_ = torch.randint(0, K, (B * P,), generator=gen_)
        lg_[torch.arange(B * P)[peak_], cls_[peak_]] += 4.0
        lg_[~peak_, K] += 4.0
        lg_, dl_ = lg_.to(dev), (torch.randn(B * P, 4 * K, generator=gen_) * 0.5).to(dev)
        pb_ = torch.cat([wl.resident["boxes"][i][:P] for i in range(B)], 0).float().contiguous()
        offs_ = torch.arange(0, B * P + 1, P, dtype=torch.int32, device=dev)
        hw_ = ops.image_hw_tensor([(H_IMG, W_IMG)] * B, dev)
        tn_, tt_ = [], []
        for i in range(6):
            flush.fill_(i)
            _lib.PROFILE = {}
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record()
            det_ = ops.fast_rcnn_inference_device(lg_, dl_, pb_, offs_, hw_, 0.05, 0.5, 100, max_rois_per_image=P)
            e1_.record()
            torch.cuda.synchronize()
            if i >= 2:
                tt_.append(e0_.elapsed_time(e1_))
                tn_.append(sum(a.elapsed_time(b) for a, b, _ in _lib.PROFILE.get("b200_batched_nms", [])))
            _lib.PROFILE = None
        nms_op = {"nms_us_per_image": 1e3 * float(np.median(tn_)) / B, "postprocess_us_per_image": 1e3 * float(np.median(tt_)) / B,
                  "candidates_per_image": float(det_["n_candidates"].float().mean()),
                  "detections_per_image": float(det_["counts"].float().mean()), "images": B, "classes": K,
                  "note": "b200_batched_nms (3 kernels: class sort, per-class NMS, merge) / whole fast_rcnn_inference on the device, "
                          "CUDA events, median of 4, bit-exact keep indices (tests/test_gpu_detect_post.py)"}
        del lg_, dl_, det_

    e2e_steps = max(3, min(args.steps, 10))
    e2e, h2d_bytes, d2h_bytes = wl.e2e(e2e_steps)
    losses_last = [float(v) for v in wl.out["losses"].tolist()] if train else None
    n_det = float(np.mean([len(r) for r in wl.out])) if not train else None
    wl.close()

    # ---- the other BASELINE configurations, briefly, in the same process ---------------------------------------------
    extra = {}
    if not args.no_extras:
        plan = [("infer_voc20", "infer", 20, False), ("infer_coco80", "infer", 80, False), ("distill_voc20", "train", 20, True)]
        if not train:
            plan = [("train_voc20", "train", 20, False)] + plan[1:]
        for name, mode, k, distill in plan:
            try:
                w2 = Workload(args, mode, k, distill, dev, rank, world)
                w2.prepare(3)
                n2 = 10
                tot2, _, l2, _, ev2 = w2.timed(n2, flush)
                st2 = w2.stage_split(flush, ev2, n=3)
                ent = {"images_per_sec": world * B * n2 / (tot2 * 1e-3), "ms_per_step": tot2 / n2, "stage_ms": st2, "gpu_launches": l2,
                       "config": workload_config(args, mode, k, distill, B)["workload"], "steps": n2, "warmup": 3,
                       "cuda_graph": w2.graph_note}
                if mode == "infer":
                    ent["nms_us_per_image"] = 1e3 * st2["decode_nms"] / B
                    ent["detections_per_image"] = float(np.mean([len(r) for r in w2.out]))
                    if world > 1:        # the evaluation exchange (pascal_voc_evaluation.py:84): padded all-gather over NCCL
                        cnt, dets = bdist.pack_detections(w2.out)
                        tg = []
                        for _ in range(4):
                            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0_.record()
                            allc, alld = bdist.all_gather_detections(cnt, dets, B * world)
                            e1_.record()
                            torch.cuda.synchronize()
                            tg.append(e0_.elapsed_time(e1_))
                        ent["all_gather_detections_ms"] = float(np.median(tg[1:]))
                        ent["all_gather_detections_images"] = int(allc.shape[0])
                extra[name] = ent
                w2.close()
                del w2
            except Exception as ex:  # noqa: BLE001
                extra[name] = {"error": repr(ex)[:300]}
            torch.cuda.synchronize()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_burst = float(peaks.get("bf16_tflops", 1590.0))
        tc_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
        R = B * P
        cfg_key = "%s_b%d_p%d_k%d" % (args.mode + ("_distill" if args.distill else ""), B, P, K)
        roi_ms = prof.get("b200_roi_align_fwd", {}).get("ms_per_step", float("nan"))
        roi_gbs = live_bytes / (roi_ms * 1e-3) / 1e9
        traffic = traffic_from_profiles(cfg_key) or {}
        roi_roof = {"kernel": "roi_slice_prepare_kernel + roi_align_fwd_slice_kernel<4,4,2> (bf16, rank 0)",
                    "bound": "hbm", "achieved": roi_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": roi_gbs / hbm_peak,
                    "traffic": traffic.get("roi_align_fwd"),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "algorithmic_bytes_per_launch": live_bytes, "avg_launch_ms": roi_ms,
                    "bins_pooled": "4x4 of 7x7 (dead bins skipped: res5 block 0 reads [::2, ::2] only)",
                    "timing": "CUDA events recorded around the entry point on the launching stream, median over %d profiled steps" % prof_steps,
                    "launched_alone_ms": roi16_ms,
                    "standalone_op": dict(roi_op, fwd_frac=roi_op["fwd_gbs"] / hbm_peak, bwd_frac=roi_op["bwd_gbs"] / hbm_peak)}
        # dominant hand-written kernel of the step = the tcgen05 GEMMs (res5 convolutions + text-fusion chain)
        gem = {"ms": 0.0, "flop": 0.0, "calls": 0.0}
        for name in ("b200_gemm_bf16", "b200_gemm_bf16_ex", "b200_gemm2"):
            if name in prof:
                gem["ms"] += prof[name]["ms_per_step"]
                gem["flop"] += prof[name].get("flop_per_step", 0.0)
                gem["calls"] += prof[name]["calls_per_step"]
        gemm_tf = gem["flop"] / (gem["ms"] * 1e-3) / 1e12 if gem["ms"] else float("nan")
        b2b_ok = b2b_ms == b2b_ms and b2b_ms > 0
        b2b_tf = gemm_flop / (b2b_ms * 1e-3) / 1e12 if b2b_ok else gemm_tf
        g2 = prof.get("b200_gemm2", {})
        gemm_roof = {"kernel": "gemm2_pair_kernel<BN,F> (tcgen05.mma.cta_group::2: res5 convolutions fwd + dgrad) + gemm_bf16_tcgen05_kernel<BN> "
                               "(text-fusion chain): all %d tensor-core launches of the step, rank 0" % round(gem["calls"]),
                     "bound": "tensor", "achieved": b2b_tf, "peak": tc_burst, "unit": "TFLOP/s", "frac": b2b_tf / tc_burst,
                     "frac_of_sustained_peak": b2b_tf / tc_sust,
                     "traffic": traffic.get("gemm"),
                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; bf16_tflops_sustained = %.0f beside it)" % tc_sust) if peaks else "fallback",
                     "algorithmic_flop_per_step": gem["flop"], "ms_per_step": b2b_ms if b2b_ok else gem["ms"],
                     "timing": "every tensor-core GEMM launch of one step (same shapes / layouts / epilogues, fresh operands) issued back to back on "
                               "the launching stream, ONE CUDA event pair around the sequence, L2 flushed before it, median of 4: the sum of the "
                               "kernels' durations.  Inside the step the weight-gradient GEMMs run on a side stream by design: see in_step",
                     "in_step": {"achieved": gemm_tf, "ms_per_step": gem["ms"], "frac": gemm_tf / tc_burst,
                                 "note": "CUDA events around every GEMM entry-point call on its launching stream, summed per step, median over "
                                         "%d profiled eager steps (SMs shared with the other streams' kernels)" % prof_steps},
                     "res5_convolutions": {"ms_per_step": g2.get("ms_per_step"), "tflops": g2.get("tflops"), "calls_per_step": g2.get("calls_per_step")}}
        ours_ms = sum(v["ms_per_step"] for v in prof.values())
        line = {
            "metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            # identical to the reference arm's `config` (same workload); how this arm executes it is beside it
            "config": workload_config(args, args.mode, K, train and args.distill, B), "cuda_graph": wl.graph_note,
            "e2e": {"value": world * B * e2e_steps / (e2e["overlap"] * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "pipeline": "double-buffered device inputs: upload of step i+1 on a copy stream overlaps compute of step i",
                    "serial_value": world * B * e2e_steps / (e2e["serial"] * 1e-3)},
            "gpu_launches": int(launches),
            "library_kernels": "none on the timed path (res5 runs on gemm2_pair_kernel; RES5_IMPL=cudnn restores the library convolutions)",
            "clocks": sampler.summary(),
            "roofline": gemm_roof, "roofline_other": roi_roof,
            "stage_ms": stage_ms,
            "stage_ms_note": "eager steps with CUDA events at the stage boundaries inside forward() (the timed steps replay one CUDA graph)" if wl.graph_note.startswith(("whole", "forward")) else "timed steps",
            "roi_align_gbs": {"in_step_4x4_bins": roi_gbs, "operator_7x7_fwd": roi_op.get("fwd_gbs"), "operator_7x7_bwd": roi_op.get("bwd_gbs")},
            "nms": nms_op,
            "own_kernels_ms_per_step": ours_ms, "host_enqueue_ms_per_step": host_ms,
            "own_kernels_profile": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "flop_per_step"}
                                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])},
            "extra": extra,
        }
        if train:
            line["losses_last_step"] = losses_last
        else:
            line["nms_us_per_image"] = 1e3 * stage_ms["decode_nms"] / B
            line["detections_per_image"] = n_det
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only: at N > 1 the other ranks would spin on the barrier meanwhile
            try:
                line["cpu_baseline"] = run_cpu_baseline(args)
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        # drop the captured graphs (they may hold NCCL work) before the communicator goes away
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
