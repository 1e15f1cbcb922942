#!/usr/bin/env python
"""bench.py — ROI-head images/s on B200 (BASELINE.json metric), with roofline, CPU baseline and e2e figures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode train|infer]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload = BASELINE.json configs[1]: "VOC split1 10-shot novel fine-tune with CLIP text-fused ROI head, bf16,
1 B200".  A "step" is one fine-tune step of the text-fused C4 ROI head over one batch of synthetic input on each GPU
(8 images x 512 sampled, labelled proposals each; 600x800 px -> res4 38x50x1024; K=20; CLIP 512-d):
  forward   GDL + affine_rcnn on the res4 map -> ROIAlign 7x7 -> res5 (frozen, FrozenBN folded, cuDNN bf16) -> mean
            -> text-fusion chain (tcgen05 GEMMs) -> cls_score (dropout 0.8) / bbox_pred -> loss_cls, loss_box_reg,
            loss_attentive
  backward  through all of it down to d(res4 map) (GDL scale 0.001) and every trained parameter (attention,
            predictor, affine_rcnn); res5 is frozen (ROI_HEADS.FREEZE_FEAT) so it only propagates the data gradient
  update    gradient all-reduce over NCCL when N > 1, then SGD + momentum on the trained parameters
`--mode infer` times the inference direction instead (forward + softmax/decode/threshold/per-class NMS/top-100).
Images shard across GPUs (weak scaling); the only exchanges are the gradient all-reduce (train) / the detection
all-gather after the loop (infer).  Proposal sampling/labelling (marked "next") is not in the step:
the synthetic proposals arrive sampled and labelled.
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
    ap.add_argument("--props", type=int, default=512)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--cpu-baseline-images", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--distill", action="store_true",
                    help="train mode: BASELINE configs[3] — the student step with the KL loss against a frozen VKV teacher "
                         "(SematicRes5ROIHeadsDistill); not the default metric line")
    ap.add_argument("--no-graph", action="store_true",
                    help="fine-tune mode: enqueue every step from the host instead of replaying one captured CUDA graph")
    ap.add_argument("--full-bins", action="store_true",
                    help="pool all 49 bins (the stand-alone ROIAlign op) instead of only the 16 that res5's stride-2 1x1 convs read")
    return ap.parse_args()


def synth_inputs(n_images, props, seed0=1234, num_classes=20):
    """SURVEY.md §8(d) synthetic inputs: post-ReLU res4 maps, RPN-like + jittered proposals, per-image seeds; for the
    fine-tune direction the proposals come sampled and labelled (25 % foreground, GT = jittered proposal)."""
    from fewshotobjectdetection_imporove_via_text_feature_b200.utils.synthetic import synth_proposals
    feat = torch.relu(torch.randn(n_images, C4, HF, WF, generator=torch.Generator().manual_seed(0)))
    boxes, gt_cls, gt_boxes = [], [], []
    for i in range(n_images):
        gen = torch.Generator().manual_seed(seed0 + i)
        b = synth_proposals(props, H_IMG, W_IMG, gen, n_obj=8)[0]
        c = torch.randint(0, num_classes, (props,), generator=gen)
        c[props // 4:] = num_classes
        g = b + torch.randn(props, 4, generator=gen) * 4
        g[:, 2:] = torch.maximum(g[:, 2:], g[:, :2] + 2)
        boxes.append(b)
        gt_cls.append(c)
        gt_boxes.append(g)
    return feat, boxes, gt_cls, gt_boxes


def build_head(num_classes, device, train=False, distill=False):
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
        _, head, aff = build_head(args.classes, "cpu", train=self.train)
        self.params = {k: v.detach().float().clone() for k, v in head.state_dict().items()}
        self.text = torch.cat([head.attention.embed, head.attention.bg_feature], 0).float()
        self.aff_w, self.aff_b = aff.weight.detach().clone(), aff.bias.detach().clone()
        self.feat, self.boxes, self.gt_cls, self.gt_boxes = synth_inputs(1, args.props, num_classes=args.classes)
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
        out = O.head_train_step(self.feat, self.boxes, self.gt_cls[0], self.gt_boxes[0], self.text, self.params, self.aff_w,
                                self.aff_b, GDL_LAMBDA, a.classes, DROP_P, stages=stages)
        self.opt.step()
        return out

    def describe(self, n):
        return "%d image(s) x %d proposals, %s step, fp32, torch %d threads" % (
            n, self.args.props, "fine-tune (fwd + bwd + SGD)" if self.train else "inference", torch.get_num_threads())


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
        "config": workload_config(args, 1), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, images_per_gpu):
    if args.mode == "train" and getattr(args, "distill", False):
        what = ("distillation fine-tune step (BASELINE configs[3]): the configs[1] step + frozen LV_attention_VKV teacher forward "
                "(GT-conditioned, GloVe 300-d) + loss_kl (T = 5) on the student's logits")
    elif args.mode == "train":
        what = ("fine-tune step (BASELINE configs[1]): GDL+affine_rcnn -> ROIAlign 7x7 -> res5 (frozen) -> text fusion -> "
                "cls_score(dropout 0.8)/bbox_pred -> loss_cls+loss_box_reg+loss_attentive -> backward to the res4 map "
                "and all trained parameters -> (grad all-reduce) -> SGD+momentum; proposals arrive sampled and labelled")
    else:
        what = "inference step: affine_rcnn -> ROIAlign 7x7 -> res5 -> text fusion -> decode/NMS top-100"
    return {"workload": "DeFRCN R-101 C4 text-fused ROI head (SematicRes5ROIHeads, CLIP 512-d, K=%d), %s" % (args.classes, what),
            "mode": args.mode, "classes": args.classes,
            "roi_align_bins": "all 49" if getattr(args, "full_bins", False) else "16 live of 49 (stride-2 consumer)",
            "images_per_gpu_per_step": images_per_gpu, "proposals_per_image": args.props, "image_px": [H_IMG, W_IMG],
            "res4_map": [C4, HF, WF], "l2": "flushed between timed steps (256 MiB write)", "parallelism": "image-sharded dp%d" % args.gpus}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank, world)

    import torch.distributed as dist
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, distributed as bdist, ops, train_ops
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads.fast_rcnn import FastRCNNOutputs
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NB the GPU box exports NCCL_DEBUG=VERSION: NCCL itself prints one "NCCL version ..." line to stdout before the JSON line
        dist.init_process_group("nccl", device_id=dev)
    train = args.mode == "train"
    B, P, K = args.images_per_gpu, args.props, args.classes
    distill = train and getattr(args, "distill", False)
    cfg, head, aff = build_head(K, dev, train=train, distill=distill)
    feat_h, boxes_h, cls_h, gtb_h = synth_inputs(B, P, seed0=1234 + 1000 * rank, num_classes=K)
    host = {"feat": feat_h.pin_memory(), "boxes": torch.stack(boxes_h).pin_memory()}
    if train:
        host["gt_cls"] = torch.stack(cls_h).pin_memory()
        host["gt_boxes"] = torch.stack(gtb_h).pin_memory()
    names = list(host)
    resident = {k: v.to(dev) for k, v in host.items()}
    sizes = [(H_IMG, W_IMG)] * B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # res5's first block reads the pooled 7x7 map through 1x1 stride-2 convs: only bins [::2, ::2] are live
    skip = head.skip_dead_bins and head.res5[0].reads_strided_1x1() and not args.full_bins
    bin_step = head.res5[0].stride if skip else 1
    nb = -(-7 // bin_step)
    opt = None
    if train:
        opt = train_ops.FlatSGD(list(head.attention.parameters()) + list(head.box_predictor.parameters()) + list(aff.parameters()),
                                lr=LR, momentum=MOMENTUM, weight_decay=WD, direct_grads=True)

    def make_props(d):
        props = []
        for i in range(B):
            inst = Instances(sizes[i])
            inst.proposal_boxes = Boxes(d["boxes"][i])
            if train:
                inst.gt_boxes = Boxes(d["gt_boxes"][i])
                inst.gt_classes = d["gt_cls"][i]
            props.append(inst)
        return props

    if train:
        stage_names = ["gdl_affine", "roi_align", "res5_mean", "text_fusion_losses", "bwd_text_fusion", "bwd_res5",
                       "bwd_roi_align", "bwd_gdl_affine", "allreduce_sgd"]
    else:
        stage_names = ["affine", "roi_align", "res5_mean", "text_fusion_predictor", "decode_nms"]
    n_marks = len(stage_names) + 1

    def step(d, ev=None, exchange=True):
        def mark(i):
            if ev is not None:
                ev[i].record()
        props = make_props(d)
        if not train:
            mark(0)
            f = aff(d["feat"], None, True, torch.bfloat16)                                    # G2 (+layout/dtype for the gather)
            mark(1)
            pooled = head.pooler([f], [p.proposal_boxes for p in props], bin_step=bin_step)   # P1
            mark(2)
            fp = head._res5_mean(pooled, prestrided=bin_step > 1)                               # P2 (cuDNN + own mean)
            mark(3)
            att, _ = head.forward_att(fp)                                                     # T1, A1-A6, C1
            mark(4)
            outs = FastRCNNOutputs(head.box2box_transform, att["pred_logits"], att["pred_bbox"], props, 0.0)
            out = outs.inference_device(head.test_score_thresh, head.test_nms_thresh, head.test_detections_per_img)  # D1-D3
            mark(5)
            return out
        mark(0)
        opt.zero_grad()
        begin = torch.cuda.Event()
        begin.record()                                                                        # parameters hold this step's values
        x = d["feat"].detach().requires_grad_(True)
        f = aff(x, GDL_LAMBDA, True, torch.bfloat16)                                          # G1 + G2
        mark(1)
        pooled = head.pooler([f], [p.proposal_boxes for p in props], bin_step=bin_step)       # P1
        mark(2)
        fp = head._res5_mean(pooled, prestrided=bin_step > 1)                                 # P2 (frozen: one node)
        head.prefetch_text_side(after=begin)                                                  # T1 (+ text half of A1/A2): side stream, under res5
        mark(3)
        gt = d["gt_cls"].reshape(-1)
        if distill:            # frozen teacher forward (fused, no autograd) + KL as the fourth loss of the fused node
            losses, _ = head.fused_train_losses(fp, props, gt, head._teacher_logits(fp, gt), head._kd_params())
        else:
            losses, _ = head.fused_train_losses(fp, props, gt)                                # T1, A1-A6, C1, L1
        mark(4)
        if ev is not None:     # events inside the backward pass: recorded when the gradient of that tensor is ready
            fp.register_hook(lambda g: ev[5].record())
            pooled.register_hook(lambda g: ev[6].record())
            f.register_hook(lambda g: ev[7].record())
        total = losses["loss_cls"] + losses["loss_box_reg"] + losses["loss_attentive"]
        if distill:
            total = total + losses["loss_kl"]
        total.backward()                                                                      # L1, A*, P2, P1b, G1/G2 bwd
        mark(8)
        res = {"losses": torch.stack([losses["loss_cls"], losses["loss_box_reg"], losses["loss_attentive"]]).detach(),
               "grad_feat": x.grad}
        if not exchange:
            opt.sync_grads()         # join the parameter-gradient streams (a captured region must end on one stream)
            return res
        if world > 1:
            # the one real exchange step: head gradients on a communication stream behind the parameter-gradient streams
            # (under the res5 / ROIAlign backward still queued on the GPU), affine_rcnn's two vectors at the end
            opt.all_reduce_grads(n_late_params=len(list(aff.parameters())))
        opt.step()
        mark(9)
        return res

    grad_ctx = torch.enable_grad() if train else torch.no_grad()
    with grad_ctx:
        for _ in range(max(args.warmup, 3)):
            out = step(resident)
        torch.cuda.synchronize()
        # ---- fine-tune step as one CUDA graph ------------------------------------------------------------
        # ~200 launches on five streams per step cost the host about as long to enqueue as the GPU takes to run them;
        # a busy host then stalls the GPU.  The whole step (zero_grad .. backward [.. SGD at world 1]) is captured once
        # and replayed; inputs are copied into the graph's static buffers, the classifier-dropout step counter lives
        # in device memory so every replay draws a new mask.  World > 1: the gradient all-reduce and SGD stay outside.
        graph, in_graph, graph_note = None, False, "off (--no-graph)" if train else "n/a"
        if train and not args.no_graph:
            try:
                head.use_device_dropout_counter(True)
                # world > 1: the NCCL gradient all-reduce is captured too (on the communication stream, behind the
                # parameter-gradient streams and under res5's backward); BENCH_GRAPH_ALLREDUCE=0 keeps it and SGD outside
                in_graph = world == 1 or os.environ.get("BENCH_GRAPH_ALLREDUCE", "1") == "1"
                graph = train_ops.GraphedStep(lambda d: step(d, exchange=in_graph), resident)
                graph_note = "whole step" if in_graph else "forward + backward (all-reduce and SGD outside)"
            except Exception as e:  # noqa: BLE001
                graph, graph_note = None, "capture failed, running eagerly: %s" % str(e).splitlines()[0][:200]
                head.use_device_dropout_counter(False)
                torch.cuda.synchronize()

        def run_step(d, ev=None):
            if graph is None:
                return step(d, ev)
            if ev is not None:
                ev[0].record()
            static_out = graph(d)
            if world > 1 and not in_graph:
                dist.all_reduce(opt.grad, op=dist.ReduceOp.AVG)
                opt.step()
            if ev is not None:
                ev[n_marks - 1].record()
            return static_out
        # ---- device-resident timing --------------------------------------------------------------------
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_marks)] for _ in range(args.steps)]
        sampler = ClockSampler(local)
        sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.LAUNCHES
        host_prof = None
        if os.environ.get("BENCH_HOST_PROFILE"):        # where the host time of a step goes (cProfile, stderr)
            import cProfile
            host_prof = cProfile.Profile()
            host_prof.enable()
        t_cpu0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            out = run_step(resident, evs[i])
        if host_prof is not None:
            import pstats
            host_prof.disable()
            pstats.Stats(host_prof, stream=sys.stderr).sort_stats("tottime").print_stats(45)
        cpu_enqueue_ms = (time.perf_counter() - t_cpu0) * 1e3 / max(args.steps, 1)    # host time to enqueue one step (no sync)
        if world > 1 and not train:
            insts = [Instances(sizes[0], pred_boxes=Boxes(out["boxes"][i]), scores=out["scores"][i], pred_classes=out["classes"][i]) for i in range(B)]
            cnt, dets = bdist.pack_detections(insts)
            bdist.all_gather_detections(out["counts"], dets, B * world)
        torch.cuda.synchronize()
        launches = (_lib.LAUNCHES - l0) // max(args.steps, 1)
        if graph is not None:        # replays bypass the Python entry points: count the kernels recorded into the graph
            launches = graph.kernel_launches
        sampler.stop_flag = True
        if world > 1:
            dist.barrier()
        per_step = [evs[i][0].elapsed_time(evs[i][n_marks - 1]) for i in range(args.steps)]
        if graph is not None:        # stage marks cannot sit inside the graph: a few eager steps give the stage split
            for _ in range(6):           # the eager path's allocations settle again after the capture
                step(resident)
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_marks)] for _ in range(5)]
            for i in range(5):
                flush.fill_(i)
                step(resident, evs[i])
            torch.cuda.synchronize()
        stage_ms = [float(np.median([e[s].elapsed_time(e[s + 1]) for e in evs])) for s in range(n_marks - 1)]
        total_ms = torch.tensor([float(sum(per_step))], device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms)
        # ---- per-entry-point profile (separate pass: an event pair around every C-ABI call) ---------------------
        _lib.PROFILE = {}
        prof_steps = 5
        for i in range(prof_steps):
            flush.fill_(i)
            step(resident)
        torch.cuda.synchronize()
        prof = {}
        fl_of = lambda t: (t[0] if isinstance(t, tuple) else t) or 0.0
        if os.environ.get("BENCH_DUMP_CALLS"):
            for name, rows in _lib.PROFILE.items():
                for a, b, t in rows[: len(rows) // prof_steps]:
                    ms = a.elapsed_time(b)
                    print("CALL %-34s %8.4f ms %s" % (name, ms, ("%.1f GF  %.0f TF/s %s" % (fl_of(t) / 1e9, fl_of(t) / ms / 1e9, t[1] if isinstance(t, tuple) else "")) if t else ""), file=sys.stderr)
        # the step's GEMM launches, each shape / epilogue re-issued ALONE on the launching stream (fresh operands of the
        # same shape, L2-warm as inside the step): inside the step the side-stream GEMMs share the SMs with res5's
        # kernels by design, so their event-bracketed times there do not measure the kernel
        gemm_cases = [t[1] for name in ("b200_gemm_bf16", "b200_gemm_bf16_ex") for _, _, t in _lib.PROFILE.get(name, [])[: len(_lib.PROFILE.get(name, [])) // prof_steps]
                      if isinstance(t, tuple)]
        for name, rows in _lib.PROFILE.items():
            per = max(len(rows) // prof_steps, 1)            # the same calls every step: median over the profiled steps
            ms = float(np.median([sum(a.elapsed_time(b) for a, b, _ in rows[i * per:(i + 1) * per]) for i in range(prof_steps)]))
            fl = sum(fl_of(t) for _, _, t in rows if t) / prof_steps
            prof[name] = {"ms_per_step": ms, "calls_per_step": len(rows) / prof_steps}
            if fl:
                prof[name]["tflops"] = fl / (ms * 1e-3) / 1e12
                prof[name]["flop_per_step"] = fl
        _lib.PROFILE = None
        # the stand-alone ROIAlign operator (all 49 bins, what torchvision.ops.roi_align computes) on the same maps / ROIs, and
        # its backward (lists planned ahead, as in the step), L2 flushed before every launch
        roi_op = {}
        with torch.enable_grad():
            fmap = resident["feat"].to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            rois_l = [resident["boxes"][i] for i in range(B)]
            rr, oo = ops.boxes_to_rois(rois_l)
            tf, tb = [], []
            for i in range(5):
                flush.fill_(i)
                e0_, e1_, e2_, e3_ = (torch.cuda.Event(enable_timing=True) for _ in range(4))
                e0_.record()
                pooled_full = ops.roi_align(fmap, rr, 7, 1.0 / 16, 0, True, channels_last_out=True, roi_batch_offsets=oo, bin_step=1)
                e1_.record()
                gfull = torch.ones_like(pooled_full)
                flush.fill_(i + 1)
                e2_.record()
                pooled_full.backward(gfull)
                e3_.record()
                torch.cuda.synchronize()
                fmap.grad = None
                if i >= 2:
                    tf.append(e0_.elapsed_time(e1_))
                    tb.append(e2_.elapsed_time(e3_))
            full_bytes = B * C4 * HF * WF * 2 + B * P * 20 + B * P * C4 * 49 * 2
            roi_op = {"bins": "7x7", "algorithmic_bytes": full_bytes,
                      "fwd_ms": float(np.median(tf)), "fwd_gbs": full_bytes / float(np.median(tf)) / 1e6,
                      "bwd_ms": float(np.median(tb)), "bwd_gbs": full_bytes / float(np.median(tb)) / 1e6,
                      "note": "b200_roi_align_fwd / b200_roi_align_bwd_planned entry points, CUDA events, median of 3"}
            del pooled_full, gfull, fmap
        # the post-processing operator (BASELINE metric "NMS us"): softmax + decode + threshold compaction, per-class NMS,
        # top-100 gather on this batch's proposals with SURVEY 8(d)'s logits (30 % of the ROIs peaked on a random foreground
        # class, the rest on background), L2 flushed before every call
        nms_op = {}
        with torch.no_grad():
            gen_ = torch.Generator().manual_seed(99)
            lg_ = torch.randn(B * P, K + 1, generator=gen_)
            peak_ = torch.rand(B * P, generator=gen_) < 0.3
            cls_ = torch.randint(0, K, (B * P,), generator=gen_)
            lg_[torch.arange(B * P)[peak_], cls_[peak_]] += 4.0
            lg_[~peak_, K] += 4.0
            lg_, dl_ = lg_.to(dev), (torch.randn(B * P, 4 * K, generator=gen_) * 0.5).to(dev)
            pb_ = torch.cat([resident["boxes"][i] for i in range(B)], 0).float().contiguous()
            offs_ = torch.arange(0, B * P + 1, P, dtype=torch.int32, device=dev)
            hw_ = ops.image_hw_tensor([(H_IMG, W_IMG)] * B, dev)
            tn_, tt_ = [], []
            for i in range(6):
                flush.fill_(i)
                _lib.PROFILE = {}
                e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0_.record()
                det_ = ops.fast_rcnn_inference_device(lg_, dl_, pb_, offs_, hw_, 0.05, 0.5, 100)
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
        gemm_alone = {"ms": 0.0, "flop": 0.0, "calls": len(gemm_cases)}
        gemm_ops = []
        for (M_, N_, K_, obf, d2_, relu_, acc_, msk_, bias_) in gemm_cases:
            ld = (N_ + 7) // 8 * 8
            a_ = torch.randn(M_, K_, device=dev).to(torch.bfloat16)
            b_ = (torch.randn(N_, K_, device=dev) * 0.05).to(torch.bfloat16)
            o_ = torch.zeros(M_, ld, device=dev, dtype=torch.bfloat16 if obf else torch.float32)[:, :N_]
            o2_ = torch.empty(M_, ld, device=dev, dtype=torch.bfloat16)[:, :N_] if d2_ else None
            m_ = torch.randn(M_, ld, device=dev).to(torch.bfloat16)[:, :N_] if msk_ else None
            bi_ = torch.randn(N_, device=dev) if bias_ else None
            gemm_ops.append((a_, b_, bi_, relu_, o_, o2_, acc_, m_))
            ts_ = []
            for _ in range(4):
                e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0_.record()
                train_ops.gemm_ex(a_, b_, bi_, relu=relu_, out=o_, out2=o2_, accumulate=acc_, mask=m_)
                e1_.record()
                torch.cuda.synchronize()
                ts_.append(e0_.elapsed_time(e1_))
            gemm_alone["ms"] += float(np.median(ts_[1:]))
            gemm_alone["flop"] += 2.0 * M_ * N_ * K_
        # the same launches BACK TO BACK on the launching stream, one event pair around the whole sequence, L2 flushed
        # before it: the sum of the kernels' durations without the cross-stream SM sharing of the step (where the
        # weight-gradient GEMMs run under res5's kernels by design) and without per-launch host latency
        gemm_b2b = {"ms": float("nan"), "flop": gemm_alone["flop"]}
        if gemm_ops:
            ts_ = []
            for it in range(6):
                flush.fill_(it)
                e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0_.record()
                for (a_, b_, bi_, relu_, o_, o2_, acc_, m_) in gemm_ops:
                    train_ops.gemm_ex(a_, b_, bi_, relu=relu_, out=o_, out2=o2_, accumulate=acc_, mask=m_)
                e1_.record()
                torch.cuda.synchronize()
                ts_.append(e0_.elapsed_time(e1_))
            gemm_b2b["ms"] = float(np.median(ts_[2:]))
        del gemm_ops
        # ---- end to end: pinned host inputs -> device, result -> host, every step ------------------------------
        # The public call with HOST buffers.  Two device input buffers: the upload of step i+1 (copy stream) overlaps
        # the compute of step i; every step's inputs are copied from pinned memory and every step's result (losses /
        # detections) is read back, all inside the timed region.  `serial` is the same loop without the overlap.
        res_keys = ["losses"] if train else ["boxes", "scores", "classes", "counts"]
        res_host = {k: torch.empty_like(out[k], device="cpu").pin_memory() for k in res_keys}
        e2e_steps = max(3, min(args.steps, 10))
        cur = torch.cuda.current_stream()
        cpy = torch.cuda.Stream()
        dbuf = [{k: torch.empty_like(resident[k]) for k in names} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            b = i & 1
            with torch.cuda.stream(cpy):
                cpy.wait_event(free[b])
                for k in names:
                    dbuf[b][k].copy_(host[k], non_blocking=True)
                ready[b].record(cpy)

        def run_e2e(n, overlap):
            for ev in free:
                ev.record(cur)
            if overlap:
                upload(0)
            for i in range(n):
                b = i & 1
                if overlap:
                    if i + 1 < n:
                        upload(i + 1)
                    cur.wait_event(ready[b])
                    o = run_step(dbuf[b])
                    free[b].record(cur)
                else:
                    o = run_step({k: host[k].to(dev, non_blocking=True) for k in names})
                for k in res_host:
                    res_host[k].copy_(o[k], non_blocking=True)
                if not overlap:
                    cur.synchronize()
            cur.synchronize()

        e2e = {}
        for name, overlap in (("serial", False), ("overlap", True)):
            run_e2e(3, overlap)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            run_e2e(e2e_steps, overlap)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            e2e[name] = float(ms)
        e2e_ms = e2e["overlap"]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops_sustained", 1366.0))
        R = B * P
        e = 2
        roi_bytes = B * C4 * HF * WF * e + R * 20 + R * C4 * nb * nb * e
        roi_ms = prof.get("b200_roi_align_fwd", {}).get("ms_per_step", float("nan"))
        roi_gbs = roi_bytes / (roi_ms * 1e-3) / 1e9
        # DRAM traffic per launch from the committed `ncu --set full` capture of exactly this configuration
        # (profiles/r01_ncu_full_step_kernels.md: dram__bytes_read.sum + dram__bytes_write.sum); other shapes: not captured
        default_cfg = train and (B, P, K, bin_step) == (8, 512, 20, 2)
        roi_traffic = 143825920 if default_cfg else None
        gemm_traffic = 652731904 if default_cfg else None           # sum over the step's 26 GEMM launches
        roi_roof = {"kernel": "roi_slice_prepare_kernel + roi_align_fwd_slice_kernel<%d,%d,%d> (bf16, rank 0)" % (nb, nb, bin_step),
                    "bound": "hbm", "achieved": roi_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": roi_gbs / hbm_peak, "traffic": roi_traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "algorithmic_bytes_per_launch": roi_bytes, "avg_launch_ms": roi_ms,
                    "bins_pooled": "%dx%d of 7x7%s" % (nb, nb, " (dead bins skipped: res5 block 0 reads [::2, ::2] only)" if bin_step > 1 else ""),
                    "timing": "CUDA events recorded around the entry point on the launching stream, median over %d profiled steps" % prof_steps,
                    "standalone_op": dict(roi_op, fwd_frac=roi_op["fwd_gbs"] / hbm_peak, bwd_frac=roi_op["bwd_gbs"] / hbm_peak) if roi_op else None}
        # dominant hand-written kernel of the step = the tcgen05 GEMM (all launches of the step together)
        gem = {"ms": 0.0, "flop": 0.0, "calls": 0.0}
        for name in ("b200_gemm_bf16", "b200_gemm_bf16_ex"):
            if name in prof:
                gem["ms"] += prof[name]["ms_per_step"]
                gem["flop"] += prof[name].get("flop_per_step", 0.0)
                gem["calls"] += prof[name]["calls_per_step"]
        gemm_tf = gem["flop"] / (gem["ms"] * 1e-3) / 1e12 if gem["ms"] else float("nan")
        alone_tf = gemm_alone["flop"] / (gemm_alone["ms"] * 1e-3) / 1e12 if gemm_alone["ms"] else float("nan")
        b2b_ok = gemm_b2b["ms"] == gemm_b2b["ms"] and gemm_b2b["ms"] > 0
        b2b_tf = gemm_b2b["flop"] / (gemm_b2b["ms"] * 1e-3) / 1e12 if b2b_ok else gemm_tf
        gemm_roof = {"kernel": "gemm_bf16_tcgen05_kernel<BN> (all %d launches of the step, rank 0)" % round(gem["calls"]),
                     "bound": "tensor", "achieved": b2b_tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": b2b_tf / tc_peak, "traffic": gemm_traffic,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long sequence)" if peaks else "fallback",
                     "algorithmic_flop_per_step": gem["flop"], "ms_per_step": gemm_b2b["ms"] if b2b_ok else gem["ms"],
                     "timing": "every GEMM launch of one step (same shapes / epilogues, fresh operands) issued back to back on the launching "
                               "stream, ONE CUDA event pair around the sequence, L2 flushed before it, median of 4: the sum of the kernels' "
                               "durations.  Inside the step the weight-gradient GEMMs run on a side stream under res5's kernels by design, so "
                               "per-call event pairs there also measure that sharing: see in_step",
                     "in_step": {"achieved": gemm_tf, "ms_per_step": gem["ms"], "frac": gemm_tf / tc_peak,
                                 "note": "CUDA events around every GEMM entry-point call on its launching stream, summed per step, median over "
                                         "%d profiled eager steps (SMs shared with the other streams' kernels)" % prof_steps},
                     "launched_alone": {"achieved": alone_tf, "ms_per_step": gemm_alone["ms"],
                                        "note": "every GEMM shape / epilogue of the step re-issued alone with a synchronize between launches: "
                                                "includes the launch latency that back-to-back launches hide"}}
        ours_ms = sum(v["ms_per_step"] for v in prof.values())
        dominant_is_gemm = gem["ms"] >= roi_ms
        line = {
            "metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(args, B), cuda_graph=graph_note),
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
                    "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in res_host.values())),
                    "pipeline": "double-buffered device inputs: upload of step i+1 on a copy stream overlaps compute of step i",
                    "serial_value": world * B * e2e_steps / (e2e["serial"] * 1e-3)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": gemm_roof if dominant_is_gemm else roi_roof,
            "roofline_other": roi_roof if dominant_is_gemm else gemm_roof,
            "stage_ms": dict(zip(stage_names, stage_ms)),
            "stage_ms_note": "eager steps (the timed steps replay one CUDA graph)" if graph is not None else "timed steps",
            "roi_align_gbs": {"in_step_%dx%d_bins" % (nb, nb): roi_gbs, "operator_7x7_fwd": roi_op.get("fwd_gbs"), "operator_7x7_bwd": roi_op.get("bwd_gbs")},
            "nms": nms_op,
            "own_kernels_ms_per_step": ours_ms, "host_enqueue_ms_per_step": cpu_enqueue_ms,
            "own_kernels_profile": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "flop_per_step"}
                                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])},
        }
        if train:
            line["losses_last_step"] = [float(v) for v in out["losses"].tolist()]
        else:
            line["nms_us_per_image"] = 1e3 * stage_ms[4] / B
            line["candidates_per_image"] = out["n_candidates"].float().mean().item()
            line["detections_per_image"] = out["counts"].float().mean().item()
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = run_cpu_baseline(args)
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        # drop the captured graph (it may hold NCCL work) before the communicator goes away
        graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
