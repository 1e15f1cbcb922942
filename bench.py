#!/usr/bin/env python
"""bench.py — ROI-head images/s on B200 (BASELINE.json metric), with roofline, CPU baseline and e2e figures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one inference pass of the text-fused C4 ROI head over one batch of synthetic input on each GPU:
  affine_rcnn(GDL) on the res4 map -> ROIAlign 7x7 -> res5 (cuDNN bf16) -> spatial mean -> text-fusion chain
  (tcgen05 GEMMs) -> cls_score / bbox_pred -> softmax + decode + threshold + per-class NMS + top-100.
Workload = BASELINE.json configs[1] shape (VOC, K=20, CLIP 512-d, 600x800 px -> res4 38x50x1024, 512
proposals per image, bf16), inference direction.  Images shard across GPUs with no data-path collective
(weak scaling); the only exchange is the detection all-gather after the loop.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images-per-gpu", type=int, default=8)
    ap.add_argument("--props", type=int, default=512)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--cpu-baseline-images", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-bins", action="store_true",
                    help="pool all 49 bins (the stand-alone ROIAlign op) instead of only the 16 that res5's stride-2 1x1 convs read")
    return ap.parse_args()


def synth_inputs(n_images, props, seed0=1234):
    """SURVEY.md §8(d) synthetic inputs: post-ReLU res4 maps, RPN-like + jittered proposals, per-image seeds."""
    from oracle.gen_golden import synth_proposals
    feat = torch.relu(torch.randn(n_images, C4, HF, WF, generator=torch.Generator().manual_seed(0)))
    boxes = [synth_proposals(props, H_IMG, W_IMG, torch.Generator().manual_seed(seed0 + i), n_obj=8)[0]
             for i in range(n_images)]
    return feat, boxes


def build_head(num_classes, device):
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config, modeling
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import ShapeSpec
    cfg = config.get_cfg()
    cfg.MODEL.ROI_HEADS.NAME = "SematicRes5ROIHeads"
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = num_classes
    cfg.MODEL.ADDITION.NAME = "clip"
    torch.manual_seed(0)
    head = modeling.build_roi_heads(cfg, {"res4": ShapeSpec(channels=C4, stride=16)}).eval()
    aff = modeling.AffineLayer(C4, bias=True)
    with torch.no_grad():
        # random-init weights of the reference architecture; classifier scaled so that scores are not uniform
        head.box_predictor.cls_score.weight.mul_(40.0)
        head.box_predictor.bbox_pred.weight.mul_(50.0)
        aff.weight.normal_(1.0, 0.05)
        aff.bias.normal_(0.0, 0.05)
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


def cpu_head_step(feat, boxes, sizes, text, params, stages=None):
    from oracle import oracle as O
    return O.head_forward(feat, boxes, sizes, text, params, stages=stages)


def cpu_reference_setup(num_classes):
    cfg, head, aff = build_head(num_classes, "cpu")
    params = {k: v.detach().float() for k, v in head.state_dict().items()}
    text = torch.cat([head.attention.embed, head.attention.bg_feature], 0).float()
    return params, text, aff


def run_cpu_baseline(n_images, props, num_classes, warm=1):
    """Oracle port (reference modules' arithmetic + torchvision CPU ops) on this box's host cores."""
    torch.set_num_threads(os.cpu_count() or 1)
    params, text, aff = cpu_reference_setup(num_classes)
    feat, boxes = synth_inputs(1, props)
    with torch.no_grad():
        f = feat * aff.weight + aff.bias
        for _ in range(warm):
            cpu_head_step(f, boxes, [(H_IMG, W_IMG)], text, params)
        stages, t0 = {}, time.perf_counter()
        for _ in range(n_images):
            f = feat * aff.weight + aff.bias
            cpu_head_step(f, boxes, [(H_IMG, W_IMG)], text, params, stages)
        dt = time.perf_counter() - t0
    return {"value": n_images / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d image(s) x %d proposals, fp32, torch %d threads, after %d warm-up" % (n_images, props, torch.get_num_threads(), warm),
            "stage_ms_per_image": {k: 1e3 * v / n_images for k, v in stages.items()}}


def main_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    params, text, aff = cpu_reference_setup(args.classes)
    feat, boxes = synth_inputs(1, args.props)
    with torch.no_grad():
        for _ in range(args.warmup):
            cpu_head_step(feat * aff.weight + aff.bias, boxes, [(H_IMG, W_IMG)], text, params)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_head_step(feat * aff.weight + aff.bias, boxes, [(H_IMG, W_IMG)], text, params)
        dt = time.perf_counter() - t0
    v = args.steps / dt
    cb = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
          "sample": "each step = 1 image x %d proposals of the same workload (bounded sample), fp32, %d torch threads" % (args.props, torch.get_num_threads())}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, images_per_gpu):
    return {"workload": "DeFRCN R-101 C4 text-fused ROI head (SematicRes5ROIHeads, CLIP 512-d, K=%d), inference step: "
                        "affine_rcnn -> ROIAlign 7x7 -> res5 -> text fusion -> decode/NMS top-100" % args.classes,
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
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib, distributed as bdist, ops
    from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads.fast_rcnn import FastRCNNOutputs
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, P, K = args.images_per_gpu, args.props, args.classes
    cfg, head, aff = build_head(K, dev)
    feat_h, boxes_h = synth_inputs(B, P, seed0=1234 + 1000 * rank)
    feat_pin = feat_h.pin_memory()
    boxes_pin = torch.stack(boxes_h).pin_memory()                 # (B,P,4)
    feat_d = feat_pin.to(dev)
    boxes_d = boxes_pin.to(dev)
    sizes = [(H_IMG, W_IMG)] * B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stage_names = ["affine", "roi_align", "res5_mean", "text_fusion_predictor", "decode_nms"]
    # res5's first block reads the pooled 7x7 map through 1x1 stride-2 convs: only bins [::2, ::2] are live
    skip = head.skip_dead_bins and head.res5[0].reads_strided_1x1() and not args.full_bins
    bin_step = head.res5[0].stride if skip else 1
    nb = -(-7 // bin_step)

    def step(feat, boxes, ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record()
        props = []
        for i in range(B):
            inst = Instances(sizes[i])
            inst.proposal_boxes = Boxes(boxes[i])
            props.append(inst)
        mark(0)
        f = aff(feat, None, True, torch.bfloat16)                                     # G2 (+layout/dtype for the gather)
        mark(1)
        pooled = head.pooler([f], [p.proposal_boxes for p in props], bin_step=bin_step)   # P1
        mark(2)
        fp = head._res5_forward(pooled, prestrided=bin_step > 1).mean(dim=[2, 3], dtype=torch.float32)   # P2 (cuDNN)
        mark(3)
        att, _ = head.forward_att(fp)                                                 # T1, A1-A6, C1
        mark(4)
        outs = FastRCNNOutputs(head.box2box_transform, att["pred_logits"], att["pred_bbox"], props, 0.0)
        out = outs.inference_device(head.test_score_thresh, head.test_nms_thresh, head.test_detections_per_img)  # D1-D3
        mark(5)
        return out

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step(feat_d, boxes_d)
        torch.cuda.synchronize()
        # ---- device-resident timing --------------------------------------------------------------------
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
        sampler = ClockSampler(local)
        sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.LAUNCHES
        ops.KERNEL_EVENTS["roi_align_fwd"] = []
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            out = step(feat_d, boxes_d, evs[i])
        if world > 1:
            insts = [Instances(sizes[0], pred_boxes=Boxes(out["boxes"][i]), scores=out["scores"][i], pred_classes=out["classes"][i]) for i in range(B)]
            cnt, dets = bdist.pack_detections(insts)
            bdist.all_gather_detections(out["counts"], dets, B * world)
        torch.cuda.synchronize()
        launches = (_lib.LAUNCHES - l0) // max(args.steps, 1)
        roi_ms = float(np.mean([a.elapsed_time(b) for a, b in ops.KERNEL_EVENTS["roi_align_fwd"]]))
        ops.KERNEL_EVENTS.clear()
        n_cand = out["n_candidates"].float().mean().item()
        n_det = out["counts"].float().mean().item()
        sampler.stop_flag = True
        if world > 1:
            dist.barrier()
        per_step = [evs[i][0].elapsed_time(evs[i][5]) for i in range(args.steps)]
        stage_ms = [float(np.mean([evs[i][s].elapsed_time(evs[i][s + 1]) for i in range(args.steps)])) for s in range(5)]
        total_ms = torch.tensor([float(sum(per_step))], device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms)
        # ---- end to end: pinned host inputs -> device, detections -> host, every step --------------------
        # The public call with HOST buffers.  Two device input buffers: the upload of step i+1 (copy stream) overlaps
        # the compute of step i; every step's inputs are copied from pinned memory and every step's detections are
        # read back, all inside the timed region.  `serial` is the same loop without the overlap.
        res_host = {k: torch.empty_like(out[k], device="cpu").pin_memory() for k in ("boxes", "scores", "classes", "counts")}
        e2e_steps = max(3, min(args.steps, 10))
        cur = torch.cuda.current_stream()
        cpy = torch.cuda.Stream()
        dbuf = [(torch.empty_like(feat_d), torch.empty_like(boxes_d)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            b = i & 1
            with torch.cuda.stream(cpy):
                cpy.wait_event(free[b])
                dbuf[b][0].copy_(feat_pin, non_blocking=True)
                dbuf[b][1].copy_(boxes_pin, non_blocking=True)
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
                    o = step(dbuf[b][0], dbuf[b][1])
                    free[b].record(cur)
                else:
                    o = step(feat_pin.to(dev, non_blocking=True), boxes_pin.to(dev, non_blocking=True))
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
        R = B * P
        e = 2
        roi_bytes = B * C4 * HF * WF * e + R * 20 + R * C4 * nb * nb * e
        roi_gbs = roi_bytes / (roi_ms * 1e-3) / 1e9
        # reference order of operations (42.5 MFLOP/ROI at K=20) vs what is executed: the d x d query
        # GEMM is folded into the cached operand Kp.Wq, so the executed count drops by 2*2048^2 per ROI
        flops_ref = R * (2 * (2048 ** 2 + 2 * 2048 * 1024 + 4096 * 2048 + 2 * 2048 * 1024) + 4 * 2048 * (K + 2) + 2 * 2048 * (5 * K + 1))
        flops_fusion = flops_ref - R * 2 * 2048 ** 2
        line = {
            "metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, B),
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(feat_pin.numel() * 4 + boxes_pin.numel() * 4),
                    "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in res_host.values())),
                    "pipeline": "double-buffered device inputs: upload of step i+1 on a copy stream overlaps compute of step i",
                    "serial_value": world * B * e2e_steps / (e2e["serial"] * 1e-3)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"kernel": "roi_slice_prepare_kernel + roi_align_fwd_slice_kernel<%d,%d,%d> (bf16, rank 0)" % (nb, nb, bin_step), "bound": "hbm", "achieved": roi_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": roi_gbs / hbm_peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "algorithmic_bytes_per_launch": roi_bytes, "avg_launch_ms": roi_ms,
                         "bins_pooled": "%dx%d of 7x7%s" % (nb, nb, " (dead bins skipped: res5 block 0 reads [::2, ::2] only)" if bin_step > 1 else ""),
                         "timing": "CUDA events recorded around the launch on the launching stream, mean over the timed steps"},
            "stage_ms": dict(zip(stage_names, stage_ms)),
            "kernels_only_images_per_sec": B / ((sum(stage_ms) - stage_ms[2]) * 1e-3),
            "fusion_chain": {"tflops_executed": flops_fusion / (stage_ms[3] * 1e-3) / 1e12,
                             "tflops_reference_equivalent": flops_ref / (stage_ms[3] * 1e-3) / 1e12,
                             "peak_tflops": peaks.get("bf16_tflops_sustained"),
                             "note": "whole text-fusion + predictor stage (8 tcgen05 GEMMs, attention core, LayerNorm, cast), event-timed; "
                                     "per-GEMM tensor-pipe figures are in profiles/"},
            "nms_us_per_image": 1e3 * stage_ms[4] / B, "candidates_per_image": n_cand, "detections_per_image": n_det,
        }
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = run_cpu_baseline(args.cpu_baseline_images, P, K)
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
