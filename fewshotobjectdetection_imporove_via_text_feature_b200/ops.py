"""Torch-facing operators over the C-ABI kernels.

PyTorch is plumbing here: device memory (caching allocator), the current CUDA stream, and autograd glue.
All arithmetic on the ROI-head hot path is done by the kernels in csrc/ through `_lib.call`.
Reference call sites are cited per op.
"""
import math
import os as _os

import torch

from . import _lib
from ._lib import BF16, F32, NCHW, NHWC


# bumped by optimizers that update parameters in place through a kernel (train_ops.FlatSGD): weight caches key on it
PARAM_GENERATION = [0]

# bench.py sets KERNEL_EVENTS = {"roi_align_fwd": []} to have CUDA events recorded right around that launch
KERNEL_EVENTS = {}


def _stream():
    """Raw handle of the current CUDA stream of the current device.  The C-level accessors cost well under a microsecond;
    `torch.cuda.current_stream()` builds a Stream object and re-validates the device on every call (> 10 us, and it is
    needed ~100 times per fine-tune step)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.B200Error("b200roi ops run on CUDA tensors only (got %s); there is no CPU fallback" % t.device)


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise _lib.B200Error("unsupported dtype %s (float32 | bfloat16)" % t.dtype)


def _layout4(t):
    """(layout flag, tensor usable as-is) for a 4-d tensor: NCHW-contiguous or channels_last."""
    if t.is_contiguous():
        return NCHW, t
    if t.is_contiguous(memory_format=torch.channels_last):
        return NHWC, t
    return NCHW, t.contiguous()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _empty4(N, C, H, W, dtype, device, channels_last):
    return torch.empty((N, C, H, W), dtype=dtype, device=device,
                       memory_format=torch.channels_last if channels_last else torch.contiguous_format)


# ---------------------------------------------------------------------------------------------------
# G1 + G2  (defrcn/modeling/meta_arch/gdl.py:6-38, rcnn.py:94-97)
# ---------------------------------------------------------------------------------------------------
class _GDLAffine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, lam, out_dtype, channels_last_out):
        _require_cuda(x, weight, bias)
        in_layout, x = _layout4(x)
        N, C, H, W = x.shape
        out_dtype = out_dtype or x.dtype
        y = _empty4(N, C, H, W, out_dtype, x.device, channels_last_out)
        w = None if weight is None else weight.detach().reshape(-1).float().contiguous()
        b = None if bias is None else bias.detach().reshape(-1).float().contiguous()
        _lib.call("b200_gdl_affine_fwd", x.data_ptr(), _ptr(w), _ptr(b), y.data_ptr(), N, C, H, W, _dt(x), in_layout,
                  _dt(y), NHWC if channels_last_out else NCHW, _stream())
        ctx.save_for_backward(x, w)
        ctx.meta = (lam, in_layout, channels_last_out, weight is not None, bias is not None,
                    None if weight is None else weight.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        lam, in_layout, cl_out, has_w, has_b, wshape = ctx.meta
        N, C, H, W = x.shape
        want_cl = cl_out
        gy = gy.contiguous(memory_format=torch.channels_last) if want_cl else gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _empty4(N, C, H, W, x.dtype, x.device, in_layout == NHWC)
        if has_w and ctx.needs_input_grad[1]:
            gw = torch.empty(C, dtype=torch.float32, device=x.device)
        if has_b and ctx.needs_input_grad[2]:
            gb = torch.empty(C, dtype=torch.float32, device=x.device)
        nbytes = _lib.lib().b200_gdl_affine_bwd_workspace_bytes(N, C, H, W)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.call("b200_gdl_affine_bwd", gy.data_ptr(), x.data_ptr(), _ptr(w), float(lam), _ptr(gx), _ptr(gw), _ptr(gb),
                  N, C, H, W, _dt(x), in_layout, _dt(gy), NHWC if want_cl else NCHW, ws.data_ptr(), nbytes, _stream())
        if gw is not None:
            gw = gw.reshape(wshape)
        if gb is not None:
            gb = gb.reshape(wshape)
        return gx, gw, gb, None, None, None


def gdl_scale(g, lam):
    """GDL backward on its own: g * lambda (same kernel, weight == NULL)."""
    _require_cuda(g)
    layout, g = _layout4(g)
    N, C, H, W = g.shape
    out = _empty4(N, C, H, W, g.dtype, g.device, layout == NHWC)
    _lib.call("b200_gdl_affine_bwd", g.data_ptr(), 0, 0, float(lam), out.data_ptr(), 0, 0, N, C, H, W, _dt(g), layout,
              _dt(g), layout, 0, 0, _stream())
    return out


def gdl_affine(x, weight=None, bias=None, lam=1.0, out_dtype=None, channels_last_out=False):
    """affine(decouple_layer(x, lam)): y = x*w + b forward, grad_x = g*w*lam backward, in one pass each.
    Optionally emits channels_last and/or bf16 so ROIAlign gathers without a re-layout."""
    return _GDLAffine.apply(x, weight, bias, lam, out_dtype, channels_last_out)


# ---------------------------------------------------------------------------------------------------
# P1 / P1b  (roi_heads.py:300-305,339-340 -> detectron2 ROIPooler -> torchvision.ops.roi_align)
# ---------------------------------------------------------------------------------------------------
PLAN_AHEAD = [True]      # build the ROIAlign backward's gather lists during the forward, on a side stream
_PLAN_STREAMS = {}


def _plan_stream(dev):
    key = (dev.type, dev.index)
    if key not in _PLAN_STREAMS:
        _PLAN_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _PLAN_STREAMS[key]


class _PlanBuffer:
    """Workspace of one ROIAlign backward plan, drawn from a small per-(device, size) free list instead of the caching
    allocator: a ~100 MB block that crosses streams every step would otherwise be re-requested while its cross-stream
    use is still pending, and the allocator answers that with a fresh cudaMalloc (a host stall of about a millisecond).
    Returned to the list when the autograd context that holds it goes away; reuse is ordered by the streams themselves
    (the next plan launch waits for the current stream, which has the previous consumer queued)."""
    _free = {}

    def __init__(self, device, nbytes):
        self.key = (device.type, device.index, int(nbytes))
        pool = _PlanBuffer._free.setdefault(self.key, [])
        self.buf = pool.pop() if pool else torch.empty(int(nbytes), dtype=torch.uint8, device=device)

    def __del__(self):
        try:
            pool = _PlanBuffer._free.setdefault(self.key, [])
            if len(pool) < 4:
                pool.append(self.buf)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


# 1: fork the backward plan BEFORE the pooling kernel.  Off by default: measured step time is the same within box noise
# (4.78-4.89 vs 4.82-4.88 ms) while the pooling kernel itself runs 15 % slower next to the list builders
PLAN_EARLY = [_os.environ.get("B200_PLAN_EARLY", "0") != "0"]


class _ROIAlign(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, roi_batch_offsets, output_size, spatial_scale, sampling_ratio, aligned,
                channels_last_out, bin_step):
        _require_cuda(feat, rois)
        in_layout, feat = _layout4(feat)
        rois = rois.detach().float().contiguous()
        N, C, H, W = feat.shape
        R = rois.shape[0]
        PH, PW = output_size
        out = _empty4(R, C, -(-PH // bin_step), -(-PW // bin_step), feat.dtype, feat.device, channels_last_out)
        nbytes = _lib.lib().b200_roi_align_fwd_workspace_bytes(N, C, H, W, R, _dt(feat), in_layout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=feat.device) if nbytes else None
        ctx.plan = None
        plan_ok = (PLAN_AHEAD[0] and ctx.needs_input_grad[0] and R > 0 and roi_batch_offsets is not None and
                   feat.dtype == torch.bfloat16 and channels_last_out and in_layout == NHWC)

        def launch_plan():
            # The backward's per-pixel gather lists depend only on the ROIs: build them now on a side stream so that the
            # backward pass is a single gather launch (PLAN_EARLY: forked before the pooling kernel instead of behind it).
            pbytes = _lib.lib().b200_roi_align_bwd_plan_bytes(N, C, H, W, R, PH, PW, int(bin_step))
            if pbytes:
                main, side = torch.cuda.current_stream(), _plan_stream(feat.device)
                side.wait_stream(main)
                holder = _PlanBuffer(feat.device, pbytes)
                with torch.cuda.stream(side):
                    _lib.call("b200_roi_align_bwd_plan", rois.data_ptr(), roi_batch_offsets.data_ptr(), N, C, H, W, R, PH, PW,
                              int(bin_step), float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                              holder.buf.data_ptr(), pbytes, side.cuda_stream)
                    done = torch.cuda.Event()
                    done.record(side)
                # `rois` / offsets stay referenced by ctx until the backward, which waits on `done` first
                ctx.plan = (holder, done)

        if plan_ok and PLAN_EARLY[0]:
            launch_plan()
        ev = KERNEL_EVENTS.get("roi_align_fwd") if KERNEL_EVENTS else None
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.call("b200_roi_align_fwd", feat.data_ptr(), rois.data_ptr(), _ptr(roi_batch_offsets), out.data_ptr(), N, C,
                  H, W, R, PH, PW, int(bin_step), float(spatial_scale), int(sampling_ratio), int(bool(aligned)), _dt(feat), in_layout,
                  NHWC if channels_last_out else NCHW, _ptr(ws), nbytes, _stream(),
                  launches=(in_layout == NCHW) + (2 if _slice_path(feat, roi_batch_offsets, channels_last_out, PH, PW, bin_step) else 1))
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        ctx.save_for_backward(rois, roi_batch_offsets)
        ctx.meta = (feat.shape, feat.dtype, in_layout, output_size, spatial_scale, sampling_ratio, aligned,
                    channels_last_out, bin_step)
        if plan_ok and not PLAN_EARLY[0]:
            launch_plan()
        return out

    @staticmethod
    def backward(ctx, g):
        rois, offs = ctx.saved_tensors
        shape, dtype, in_layout, (PH, PW), scale, sr, aligned, cl_out, bin_step = ctx.meta
        if offs is None:
            raise _lib.B200Error("roi_align backward needs roi_batch_offsets (ROIs grouped by image)")
        N, C, H, W = shape
        R = rois.shape[0]
        g = g.contiguous(memory_format=torch.channels_last) if cl_out else g.contiguous()
        gin = _empty4(N, C, H, W, dtype, g.device, in_layout == NHWC)
        g_layout = NHWC if cl_out else NCHW
        if ctx.plan is not None and g.dtype == torch.bfloat16 and g.data_ptr() % 16 == 0:
            holder, done = ctx.plan
            torch.cuda.current_stream().wait_event(done)
            _lib.call("b200_roi_align_bwd_planned", g.data_ptr(), holder.buf.data_ptr(), holder.buf.numel(), gin.data_ptr(), N, C,
                      H, W, R, PH, PW, int(bin_step), _stream())
            return gin, None, None, None, None, None, None, None, None
        nbytes = _lib.lib().b200_roi_align_bwd_workspace_bytes(N, C, H, W, R, PH, PW, int(bin_step), _dt(g), in_layout,
                                                               g_layout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
        _lib.call("b200_roi_align_bwd", g.data_ptr(), rois.data_ptr(), offs.data_ptr(), gin.data_ptr(), N, C, H, W, R,
                  PH, PW, int(bin_step), float(scale), int(sr), int(bool(aligned)), _dt(g), g_layout, in_layout,
                  ws.data_ptr(), nbytes, _stream(),
                  launches=(3 if C % 64 == 0 else 4) if (g.dtype == torch.bfloat16 and cl_out and in_layout == NHWC and PH <= 7 and
                                                         PW <= 7 and C % 8 == 0 and H <= 256 and W <= 256) else 2)
        # tile path: prepare, build, gather; per-pixel CSR path (C % 64 != 0): prepare, count, fill, gather
        return gin, None, None, None, None, None, None, None, None


def _slice_path(feat, offs, cl_out, PH, PW, bin_step):
    """Mirror of the dispatch in csrc/roi_align.cu (for launch accounting only): the slice-resident kernel runs
    prepare + main."""
    if offs is None or feat.dtype != torch.bfloat16 or not cl_out or (PH, PW) != (7, 7) or bin_step not in (1, 2):
        return False
    N, C, H, W = feat.shape
    Wp = (W + 7) & ~7
    return C % 32 == 0 and H <= 256 and W <= 248 and H * Wp * 64 + 16 * 2 * 1024 + 16 * 3584 + 1024 <= 227 * 1024


def roi_align(feat, rois, output_size, spatial_scale, sampling_ratio=0, aligned=True, channels_last_out=False,
              roi_batch_offsets=None, bin_step=1):
    """torchvision.ops.roi_align semantics.  feat (N,C,H,W) NCHW or channels_last, fp32|bf16; rois (R,5).
    roi_batch_offsets: int32 (N+1) prefix of per-image ROI counts (ROIs grouped by image) — required for backward
    and for the slice-resident bf16 kernel.  bin_step=s returns only the bins [::s, ::s] of the pooled map."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    return _ROIAlign.apply(feat, rois, roi_batch_offsets, tuple(output_size), spatial_scale, sampling_ratio, aligned,
                           channels_last_out, int(bin_step))


_INDEX_CACHE = {}


def _roi_index(counts, device):
    """(batch-index column (R,1) fp32, int32 offsets (N+1)) on `device`, cached per tuple of per-image ROI counts
    (they repeat every step: 512 in training, <=1000 at test) so that no host->device copy sits on the hot path."""
    key = (counts, str(device))
    hit = _INDEX_CACHE.get(key)
    if hit is None:
        if len(_INDEX_CACHE) > 256:
            _INDEX_CACHE.clear()
        c = torch.tensor(counts, dtype=torch.int64)
        idx = torch.repeat_interleave(torch.arange(len(counts), dtype=torch.float32), c)[:, None]
        offs = torch.tensor([0] + c.cumsum(0).tolist(), dtype=torch.int32)
        hit = (idx.to(device), offs.to(device))
        _INDEX_CACHE[key] = hit
    return hit


def boxes_to_rois(box_tensors):
    """detectron2 convert_boxes_to_pooler_format: list[(Ri,4)] -> (rois (R,5), int32 offsets (N+1))."""
    dev = box_tensors[0].device
    idx, offs = _roi_index(tuple(int(b.shape[0]) for b in box_tensors), dev)
    boxes = box_tensors[0] if len(box_tensors) == 1 else cat_adjacent(list(box_tensors))
    return torch.cat([idx, boxes.float()], dim=1), offs


_HW_CACHE = {}


def image_hw_tensor(image_shapes, device):
    key = (tuple((float(h), float(w)) for h, w in image_shapes), str(device))
    hit = _HW_CACHE.get(key)
    if hit is None:
        if len(_HW_CACHE) > 256:
            _HW_CACHE.clear()
        hit = torch.tensor(key[0], dtype=torch.float32).reshape(-1, 2).to(device)
        _HW_CACHE[key] = hit
    return hit


# ---------------------------------------------------------------------------------------------------
# D1..D3  (fast_rcnn.py:46-134, :306-334)
# ---------------------------------------------------------------------------------------------------
def softmax_decode_compact(scores, deltas, proposals, roi_offsets, image_hw, score_thresh, weights=(10.0, 10.0, 5.0, 5.0),
                           input_is_prob=False, want_probs=True, max_rois_per_image=None):
    """Returns dict(probs, cand_boxes, cand_scores, cand_roi, cand_cls, cand_count, seg_offsets)."""
    _require_cuda(scores, deltas, proposals)
    scores, deltas, proposals = scores.float().contiguous(), deltas.float().contiguous(), proposals.float().contiguous()
    R, K = scores.shape[0], scores.shape[1] - 1
    N = roi_offsets.numel() - 1
    agnostic = deltas.shape[1] == 4 and K != 1
    dev = scores.device
    probs = torch.empty_like(scores) if want_probs else None
    cap = max(R * K, 1)
    cb = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    cs = torch.empty(cap, dtype=torch.float32, device=dev)
    cr = torch.empty(cap, dtype=torch.int32, device=dev)
    cc = torch.empty(cap, dtype=torch.int32, device=dev)
    cnt = torch.zeros(max(N, 1), dtype=torch.int32, device=dev)
    # an upper bound of the per-image ROI count without a host read: the caller's, else the total; many CTAs per image
    mx = int(max_rois_per_image) if max_rois_per_image else R
    nb = _lib.lib().b200_softmax_decode_compact_workspace_bytes(N, mx) if (N and mx) else 0
    ws = torch.empty(max(nb, 4), dtype=torch.uint8, device=dev)
    _lib.call("b200_softmax_decode_compact", scores.data_ptr(), int(input_is_prob), deltas.data_ptr(), proposals.data_ptr(),
              roi_offsets.data_ptr(), image_hw.data_ptr(), N, R, K, int(agnostic), *map(float, weights),
              float(score_thresh), _ptr(probs), cb.data_ptr(), cs.data_ptr(), cr.data_ptr(), cc.data_ptr(),
              cnt.data_ptr(), mx if nb else 0, ws.data_ptr() if nb else 0, nb, _stream())
    return dict(probs=probs, cand_boxes=cb, cand_scores=cs, cand_roi=cr, cand_cls=cc, cand_count=cnt[:N],
                seg_offsets=(roi_offsets * K).to(torch.int32), capacity=cap)


def batched_nms_segments(boxes, scores, classes, seg_offsets, seg_count, num_classes, iou_thresh, max_keep, max_class_slice=0):
    """Device-side batched NMS over N segments; returns (keep (N,max_keep) int32 segment-relative, keep_count (N)).
    max_class_slice: host-side bound of one class's boxes in one segment (0 = unknown), see include/b200roi.h."""
    _require_cuda(boxes, scores, classes)
    N = seg_count.numel()
    cap = boxes.shape[0]
    dev = boxes.device
    keep = torch.empty((max(N, 1), max(max_keep, 1)), dtype=torch.int32, device=dev)
    kc = torch.zeros(max(N, 1), dtype=torch.int32, device=dev)
    nbytes = _lib.lib().b200_batched_nms_workspace_bytes(N, cap, num_classes)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.call("b200_batched_nms", boxes.data_ptr(), scores.data_ptr(), classes.data_ptr(), seg_offsets.data_ptr(),
              seg_count.data_ptr(), N, cap, num_classes, float(iou_thresh), int(max_keep), int(max_class_slice),
              keep.data_ptr(), kc.data_ptr(), ws.data_ptr(), nbytes, _stream())
    return keep[:N, :max_keep], kc[:N]


def batched_nms(boxes, scores, idxs, iou_threshold):
    """Drop-in for detectron2.layers.batched_nms (fast_rcnn.py:125): returns int64 keep indices sorted by score.
    The variable-length result forces one device->host read of the count."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    dev = boxes.device
    boxes, scores = boxes.float().contiguous(), scores.float().contiguous()
    cls = idxs.to(torch.int32).contiguous()
    num_classes = int(cls.max().item()) + 1
    so = torch.zeros(1, dtype=torch.int32, device=dev)
    sc = torch.full((1,), n, dtype=torch.int32, device=dev)
    keep, kc = batched_nms_segments(boxes, scores, cls, so, sc, num_classes, iou_threshold, n)
    return keep[0, :int(kc.item())].to(torch.int64)


def fast_rcnn_inference_device(scores, deltas, proposals, roi_offsets, image_hw, score_thresh, nms_thresh, topk,
                               weights=(10.0, 10.0, 5.0, 5.0), input_is_prob=False, want_probs=False,
                               max_rois_per_image=None):
    """Whole post-processing on the device, no host synchronisation.  Returns padded tensors:
    boxes (N,topk,4), scores (N,topk), classes (N,topk) int64, roi_inds (N,topk) int64, counts (N) int32,
    n_candidates (N) int32.  max_rois_per_image: host-side bound of an image's ROI count (default: the total)."""
    c = softmax_decode_compact(scores, deltas, proposals, roi_offsets, image_hw, score_thresh, weights, input_is_prob,
                               want_probs, max_rois_per_image)
    N = roi_offsets.numel() - 1
    K = scores.shape[1] - 1
    dev = scores.device
    topk = int(topk) if topk >= 0 else c["capacity"]
    # a ROI yields at most one candidate per class: a class slice is bounded by the image's ROI count
    slice_bound = int(max_rois_per_image) if max_rois_per_image else int(scores.shape[0])
    keep, kc = batched_nms_segments(c["cand_boxes"], c["cand_scores"], c["cand_cls"], c["seg_offsets"], c["cand_count"],
                                    K, nms_thresh, topk, max_class_slice=slice_bound)
    keep = keep.contiguous()
    ob = torch.empty((N, topk, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((N, topk), dtype=torch.float32, device=dev)
    oc = torch.empty((N, topk), dtype=torch.int64, device=dev)
    orr = torch.empty((N, topk), dtype=torch.int64, device=dev)
    _lib.call("b200_gather_detections", c["cand_boxes"].data_ptr(), c["cand_scores"].data_ptr(), c["cand_roi"].data_ptr(),
              c["cand_cls"].data_ptr(), c["seg_offsets"].data_ptr(), keep.data_ptr(), kc.data_ptr(), N, topk,
              ob.data_ptr(), os_.data_ptr(), oc.data_ptr(), orr.data_ptr(), _stream())
    return dict(boxes=ob, scores=os_, classes=oc, roi_inds=orr, counts=kc, n_candidates=c["cand_count"], probs=c["probs"],
                keep=keep, cand=c)


# ---------------------------------------------------------------------------------------------------
# the stages either side of the head
# ---------------------------------------------------------------------------------------------------
def rpn_select_proposals(proposals, logits, level_sizes, image_hw, nms_thresh, pre_nms_topk, post_nms_topk,
                         min_box_size=0.0):
    """find_top_rpn_proposals (proposal_generator/proposal_utils.py:13-118) on the device, no host synchronisation.
    proposals (N, A, 4) / logits (N, A) with the levels concatenated along A in `level_sizes` order.  Returns padded
    dict(boxes (N,post,4), logits (N,post), counts (N) int32, n_invalid (N) int32)."""
    _require_cuda(proposals, logits)
    proposals, logits = proposals.float().contiguous(), logits.float().contiguous()
    N, A = logits.shape
    L = len(level_sizes)
    assert sum(level_sizes) == A and proposals.shape == (N, A, 4)
    dev = logits.device
    key = (tuple(int(v) for v in level_sizes), str(dev))
    lo = _LEVEL_CACHE.get(key)
    if lo is None:
        if len(_LEVEL_CACHE) > 256:
            _LEVEL_CACHE.clear()
        acc = [0]
        for v in level_sizes:
            acc.append(acc[-1] + int(v))
        lo = torch.tensor(acc, dtype=torch.int32).to(dev)
        _LEVEL_CACHE[key] = lo
    cap = sum(min(int(pre_nms_topk), int(v)) for v in level_sizes)
    post = int(post_nms_topk)
    ob = torch.empty((N, post, 4), dtype=torch.float32, device=dev)
    ol = torch.empty((N, post), dtype=torch.float32, device=dev)
    oc = torch.zeros(max(N, 1), dtype=torch.int32, device=dev)
    bad = torch.zeros(max(N, 1), dtype=torch.int32, device=dev)
    if N == 0 or cap == 0:
        return dict(boxes=ob, logits=ol, counts=oc[:N], n_invalid=bad[:N])
    nbytes = _lib.lib().b200_rpn_select_workspace_bytes(N, cap, L, post)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.call("b200_rpn_select_proposals", proposals.data_ptr(), logits.data_ptr(), lo.data_ptr(), image_hw.data_ptr(),
              N, A, L, int(pre_nms_topk), post, cap, float(nms_thresh), float(min_box_size), ob.data_ptr(), ol.data_ptr(),
              oc.data_ptr(), bad.data_ptr(), ws.data_ptr(), nbytes, _stream())
    return dict(boxes=ob, logits=ol, counts=oc[:N], n_invalid=bad[:N])


_LEVEL_CACHE = {}


def detector_postprocess_(boxes, scores, classes, roi_inds, counts, image_sizes, output_sizes):
    """detectron2 detector_postprocess (rcnn.py:69-73) in place on padded detections: boxes (N,topk,4), scores (N,topk),
    classes / roi_inds (N,topk) int64 or None, counts (N) int32; image_sizes / output_sizes: per image (h, w)."""
    _require_cuda(boxes, scores)
    N, topk = scores.shape
    assert boxes.is_contiguous() and scores.is_contiguous() and counts.is_contiguous()
    key = (tuple((int(h), int(w)) for h, w in image_sizes), tuple((int(h), int(w)) for h, w in output_sizes), str(boxes.device))
    hit = _POST_CACHE.get(key)
    if hit is None:
        if len(_POST_CACHE) > 256:
            _POST_CACHE.clear()
        # python-float ratios like the reference, rounded to fp32 where torch multiplies an fp32 tensor by them
        sc = [(ow / w, oh / h) for (h, w), (oh, ow) in zip(key[0], key[1])]
        hit = (torch.tensor(sc, dtype=torch.float64).to(torch.float32).reshape(-1, 2).to(boxes.device),
               torch.tensor(key[1], dtype=torch.float32).reshape(-1, 2).to(boxes.device))
        _POST_CACHE[key] = hit
    _lib.call("b200_detector_postprocess", boxes.data_ptr(), scores.data_ptr(), _ptr(classes), _ptr(roi_inds),
              counts.data_ptr(), hit[0].data_ptr(), hit[1].data_ptr(), N, topk, _stream())
    return boxes, scores, classes, roi_inds, counts


_POST_CACHE = {}


# ---------------------------------------------------------------------------------------------------
# Q2  (calibration_layer.py:110-123)
# ---------------------------------------------------------------------------------------------------
def pcb_cosine_blend_(scores, feats, prototypes, classes, exclude_mask, alpha, lower, upper):
    """In-place PCB score calibration; scores (n) sorted descending, feats (n,D), prototypes (K,D),
    classes (n) int64, exclude_mask (K) uint8 or None."""
    _require_cuda(scores, feats, prototypes, classes)
    assert scores.dtype == torch.float32 and scores.is_contiguous()
    feats, prototypes = feats.float().contiguous(), prototypes.float().contiguous()
    classes = classes.to(torch.int64).contiguous()
    n, D = feats.shape
    K = prototypes.shape[0]
    _lib.call("b200_pcb_cosine_blend", scores.data_ptr(), feats.data_ptr(), prototypes.data_ptr(), classes.data_ptr(),
              _ptr(exclude_mask), n, D, K, float(alpha), float(lower), float(upper), _stream())
    return scores


# ---------------------------------------------------------------------------------------------------
# text fusion chain (attentive_modules.py:114-177,262-294; fast_rcnn.py:403-417,462-476)
# ---------------------------------------------------------------------------------------------------
def gemm_bf16(a, b, bias=None, relu=False, out=None, out_dtype=torch.float32, out2=None):
    """out[M,N] = act(a[M,K] @ b[N,K]^T + bias).  a, b bf16 with unit inner stride (row stride free)."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1 and out.shape == (M, N)
    b32 = None if bias is None else bias.detach().float().contiguous()
    _lib.call("b200_gemm_bf16", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), _ptr(b32), out.data_ptr(),
              out.stride(0), _dt(out), _ptr(out2), 0 if out2 is None else out2.stride(0), M, N, K, int(relu), _stream(),
              tag=(2.0 * M * N * K, (M, N, K, out.dtype == torch.bfloat16, out2 is not None, bool(relu), False, False,
                                     bias is not None)))
    return out


GEMM2_TILE_N = [0]          # tests / microbenchmarks force 128 or 256
GEMM2_MAX_CLUSTERS = [0]    # tests lower it to force several tiles per CTA pair on small shapes
GEMM2_GENERIC_EPILOGUE = [0]   # tests: 1 = always the generic epilogue instantiation
import os as _os
GEMM2_NO_PDL = [int(_os.environ.get("B200_GEMM2_NO_PDL", "0"))]      # 1: no programmatic dependent launch (A/B runs)
GEMM2_SPLIT_K = [int(_os.environ.get("B200_GEMM2_SPLIT_K", "0"))]   # 0 = choose per shape, 1 = never split, n > 1 = force (tests)
_SPLITK_WS = {}                # (device, stream) -> workspace: split-K launches on one stream are ordered, streams do not share


def _splitk_plan(M, N, nkb, has_res):
    """(tile_n, split_k) for a product with few output tiles and a long K: enough K slices to give every SM pair work."""
    forced = GEMM2_SPLIT_K[0]
    if has_res or forced == 1:
        return 0, 1
    mt = -(-M // 256)
    bn = 128 if (N <= 128 or mt * -(-N // 128) <= 74) else 256
    bn = GEMM2_TILE_N[0] or bn
    tiles = mt * -(-N // bn)
    if forced > 1:
        s = forced
    else:
        # only products with a short M (weight gradients of cls_score / bbox_pred, the (K+2)-row text operands): their
        # partial tiles are a few rows; a tall skinny-N product would write whole 64-column chunks per slice
        if tiles > 18 or nkb < 16 or M > 128:
            return 0, 1
        s = min(8, 74 // tiles, nkb // 4)
    while s > 1 and -(-nkb // s) * (s - 1) >= nkb:      # every slice needs at least one K block
        s -= 1
    return (bn, s) if s > 1 else (0, 1)


def _splitk_workspace(dev, nbytes):
    key = (dev.index, _stream())
    ws = _SPLITK_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        # zero-filled once: the arrival counters at the tail must start at zero; the kernel leaves them zero
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _SPLITK_WS[key] = ws
    return ws



def gemm2(a, b, *, a2=None, a_mn=False, b_mn=False, conv_c=0, bias=None, residual=None, relu=False, mask_act=None,
          mask_bits=None, out=None, out2=None, out_f32=None, accumulate=False, bits_out=None, rowmean_out=None,
          rowsumsq_out=None, row_scale_sumsq=None, row_scale_eps=1e-12, softmax=False, gate=False,
          want_out=True, M=None):
    """CTA-pair tcgen05 GEMM / implicit 3x3 convolution (csrc/gemm2_tcgen05.cu, `b200_gemm2`).

    out[M,N] = epi([a | a2] @ b^T): a (M,K) bf16 (a_mn: stored (K,M)); conv_c: `a` is the NHWC activation (R,4,4,C) viewed
    as (16 R, C) and K = 9 C; b (N,K) bf16 (b_mn: stored (K,N)).  Epilogue: + bias, + residual (bf16), ReLU, ReLU-backward
    gate (mask_bits packed uint32 | mask_act bf16).  Outputs: `out` bf16 (allocated unless want_out=False), out2 bf16,
    out_f32 (+= when accumulate), bits_out (packed out > 0), rowmean_out (mean over groups of 16 rows)."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(-1) == 1 and b.stride(-1) == 1
    if conv_c:
        assert a.dim() == 2 and a.shape[1] == conv_c and a.is_contiguous() and a.shape[0] % 16 == 0
        Mx, K = a.shape[0], 9 * conv_c
    elif a_mn:
        K, Mx = a.shape
    else:
        Mx, K = a.shape
    M = Mx if M is None else M
    K2 = 0 if a2 is None else a2.shape[1]
    if b_mn:
        assert b.shape[0] == K and a2 is None
        N = b.shape[1]
    else:
        N = b.shape[0]
        assert b.shape[1] == K + K2, (a.shape, b.shape, K2)
    dev = a.device
    if out is None and want_out:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    for t in (out, out2, residual, mask_act):
        assert t is None or (t.dtype == torch.bfloat16 and tuple(t.shape) == (M, N) and t.stride(1) == 1), (M, N)
    assert out_f32 is None or (out_f32.dtype == torch.float32 and tuple(out_f32.shape) == (M, N) and out_f32.stride(1) == 1)
    b32 = None if bias is None else bias.detach().float().contiguous()
    d = _lib.Gemm2Desc()
    d.A, d.lda = a.data_ptr(), a.stride(0)
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and a2.shape[0] == M and a2.stride(1) == 1
        d.A2, d.lda2, d.K2 = a2.data_ptr(), a2.stride(0), K2
    d.B, d.ldb = b.data_ptr(), b.stride(0)
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn, d.conv_c = int(a_mn), int(b_mn), int(conv_c)
    d.bias = _ptr(b32)
    if residual is not None:
        d.residual, d.ld_res = residual.data_ptr(), residual.stride(0)
    d.relu = int(relu)
    if mask_act is not None:
        d.mask_act, d.ld_mask = mask_act.data_ptr(), mask_act.stride(0)
    if mask_bits is not None:
        assert mask_bits.dtype in (torch.int32, torch.uint32) and mask_bits.shape[0] == M and mask_bits.stride(1) == 1
        d.mask_bits, d.ld_mask_bits = mask_bits.data_ptr(), mask_bits.stride(0)
    if out is not None:
        d.out_bf16, d.ld_out = out.data_ptr(), out.stride(0)
    if out2 is not None:
        d.out2_bf16, d.ld_out2 = out2.data_ptr(), out2.stride(0)
    if out_f32 is not None:
        d.out_f32, d.ld_out_f32, d.accumulate = out_f32.data_ptr(), out_f32.stride(0), int(accumulate)
    if bits_out is not None:
        assert bits_out.dtype in (torch.int32, torch.uint32) and tuple(bits_out.shape) == (M, N // 32) and bits_out.stride(1) == 1
        d.bits_out, d.ld_bits_out = bits_out.data_ptr(), bits_out.stride(0)
    if rowmean_out is not None:
        assert rowmean_out.dtype == torch.float32 and tuple(rowmean_out.shape) == (M // 16, N) and rowmean_out.stride(1) == 1
        d.rowmean_out, d.ld_rowmean = rowmean_out.data_ptr(), rowmean_out.stride(0)
    row_ops = rowsumsq_out is not None or row_scale_sumsq is not None or softmax or gate
    if rowsumsq_out is not None:       # (M, ceil(N/64)) fp32: squared norm of each output row, one entry per 64-column chunk
        assert rowsumsq_out.dtype == torch.float32 and tuple(rowsumsq_out.shape) == (M, -(-N // 64)) and rowsumsq_out.stride(1) == 1
        d.rowsumsq_out, d.ld_rowsumsq = rowsumsq_out.data_ptr(), rowsumsq_out.stride(0)
    if row_scale_sumsq is not None:    # acc * 1 / max(sqrt(sum of the row's entries), eps)
        assert row_scale_sumsq.dtype == torch.float32 and row_scale_sumsq.shape[0] == M and row_scale_sumsq.stride(1) == 1
        d.row_scale_sumsq, d.ld_row_scale_sumsq = row_scale_sumsq.data_ptr(), row_scale_sumsq.stride(0)
        d.row_scale_parts, d.row_scale_eps = row_scale_sumsq.shape[1], float(row_scale_eps)
    d.softmax, d.gate = int(softmax), int(gate)
    d.tile_n, d.max_clusters, d.epilogue_variant = GEMM2_TILE_N[0], GEMM2_MAX_CLUSTERS[0], GEMM2_GENERIC_EPILOGUE[0]
    d.no_pdl = GEMM2_NO_PDL[0]
    nkb = 9 * (conv_c // 64) if conv_c else -(-K // 64) + -(-K2 // 64)
    bn_, sk = (0, 1) if row_ops else _splitk_plan(M, N, nkb, residual is not None)
    if sk > 1:
        nbytes = _lib.lib().b200_gemm2_splitk_workspace_bytes(M, N, bn_, sk)
        ws = _splitk_workspace(dev, nbytes)
        # the arrival counters sit at the head of the workspace and every launch leaves them zero
        d.tile_n, d.split_k, d.splitk_workspace, d.splitk_workspace_bytes = bn_, sk, ws.data_ptr(), ws.numel()
    import ctypes
    _lib.call("b200_gemm2", ctypes.byref(d), _stream(),
              tag=(2.0 * M * N * (K + K2),
                   dict(M=M, N=N, K=K, K2=K2, conv_c=int(conv_c), a_mn=bool(a_mn), b_mn=bool(b_mn), relu=bool(relu),
                        acc=bool(accumulate), out=out is not None, out2=out2 is not None, f32=out_f32 is not None,
                        bias=bias is not None, res=residual is not None, mbits=mask_bits is not None,
                        mact=mask_act is not None, bout=bits_out is not None, mean=rowmean_out is not None,
                        ssq=rowsumsq_out is not None, rscale=None if row_scale_sumsq is None else row_scale_sumsq.shape[1],
                        softmax=bool(softmax), gate=bool(gate))))
    return out


def text_attention(q, x, kp, vp, p1, p2, scores=None):
    """softmax(q Kp^T / sqrt(d)) Vp and the gate operands P1 = O*x, P2 = x-O (bf16, written into p1/p2).
    With `scores` (R,L) fp32 given (already scaled), q/kp are not used."""
    R, d = x.shape
    L = vp.shape[0]
    attn = torch.empty((R, L), dtype=torch.float32, device=x.device)
    assert p1.stride(0) == p2.stride(0) and p1.stride(1) == 1
    if scores is not None:
        assert scores.shape == (R, L) and scores.is_contiguous() and scores.dtype == torch.float32
    _lib.call("b200_text_attention", _ptr(None if scores is not None else q), _ptr(scores), x.data_ptr(), _dt(x),
              _ptr(kp), vp.data_ptr(), attn.data_ptr(), p1.data_ptr(), p2.data_ptr(), p1.stride(0), R, d, L, _stream())
    return attn


def residual_layernorm(y, y2, gamma, beta, eps=1e-5, relu=False, want_f32=True, want_bf16=True):
    R, d = y.shape
    of = torch.empty((R, d), dtype=torch.float32, device=y.device) if want_f32 else None
    ob = torch.empty((R, d), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    _lib.call("b200_residual_layernorm", y.data_ptr(), _ptr(y2), gamma.data_ptr(), beta.data_ptr(), float(eps), int(relu),
              _ptr(of), _ptr(ob), R, d, _stream())
    return of, ob


def cast_bf16_into(src, dst):
    """fp32 (rows, cols) -> bf16 view `dst` (rows, cols) with arbitrary row stride."""
    rows, cols = src.shape
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.stride(1) == 1 and dst.stride(1) == 1
    _lib.call("b200_cast_bf16", src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, _stream())
    return dst


_LS_CALLS = [0]


def cat_adjacent(ts):
    """torch.cat(ts, 0) — as a VIEW (no kernel) when the tensors already sit back to back in one allocation, which is how
    per-image slices of a batched tensor arrive."""
    if len(ts) == 1:
        return ts[0].contiguous()
    t0 = ts[0]
    ok = all(t.is_contiguous() and t.dtype == t0.dtype and t.shape[1:] == t0.shape[1:] and t.device == t0.device for t in ts)
    if ok and t0.numel():
        try:
            same = all(t.untyped_storage().data_ptr() == t0.untyped_storage().data_ptr() for t in ts)
        except Exception:  # noqa: BLE001
            same = False
        if same:
            ptr, adj = t0.data_ptr(), True
            for t in ts:
                if t.data_ptr() != ptr:
                    adj = False
                    break
                ptr += t.numel() * t.element_size()
            if adj:
                n = sum(t.shape[0] for t in ts)
                return torch.as_strided(t0, (n,) + tuple(t0.shape[1:]), t0.stride())
    return torch.cat(ts, 0).contiguous()


def label_and_sample_proposals(prop_boxes, gt_boxes, gt_classes, num_classes, iou_thresh=0.5, batch_per_image=512,
                               positive_fraction=0.25, seed=None, want_labels=False, seed_salt=None, append_gt=False,
                               pad_background=False):
    """S1 on the device, one launch for the batch (csrc/label_sample.cu; reference roi_heads.py:157-250).
    prop_boxes / gt_boxes / gt_classes: per-image lists of (P_i,4) fp32, (M_i,4) fp32, (M_i,) int64 CUDA tensors.
    Returns dict: sampled_idx (N,B) int32, boxes (N,B,4), classes (N,B) int64, gt_boxes (N,B,4), counts (N,2) int32
    [(#fg rows, #valid rows); rows past #valid are padding with class -1], and with want_labels the per-proposal
    matched_idx / matched_label (concatenated over images)."""
    N = len(prop_boxes)
    dev = prop_boxes[0].device
    _require_cuda(*prop_boxes)
    pc, gc = [int(b.shape[0]) for b in prop_boxes], [int(b.shape[0]) for b in gt_boxes]
    props = cat_adjacent([b.detach().float().reshape(-1, 4) for b in prop_boxes])
    gts = cat_adjacent([b.detach().float().reshape(-1, 4) for b in gt_boxes]) if sum(gc) else torch.zeros((0, 4), device=dev)
    gcl = cat_adjacent([c.detach().to(torch.int64).reshape(-1) for c in gt_classes]) if sum(gc) else torch.zeros(0, dtype=torch.int64, device=dev)
    _, poff = _roi_index(tuple(pc), dev)
    _, goff = _roi_index(tuple(gc), dev)
    B = int(batch_per_image)
    out = {"sampled_idx": torch.empty((N, B), dtype=torch.int32, device=dev),
           "boxes": torch.empty((N, B, 4), dtype=torch.float32, device=dev),
           "classes": torch.empty((N, B), dtype=torch.int64, device=dev),
           "gt_boxes": torch.empty((N, B, 4), dtype=torch.float32, device=dev),
           "counts": torch.empty((N, 2), dtype=torch.int32, device=dev)}
    n_cand = sum(pc) + (sum(gc) if append_gt else 0)
    mi = torch.empty(n_cand, dtype=torch.int32, device=dev) if want_labels else None
    ml = torch.empty(n_cand, dtype=torch.int32, device=dev) if want_labels else None
    if seed is None:
        if seed_salt is None:
            _LS_CALLS[0] += 1
        seed = (torch.initial_seed() * 2654435761 + (_LS_CALLS[0] if seed_salt is None else 0)) & 0x7FFFFFFFFFFFFFFF
    _lib.call("b200_label_sample_proposals", props.data_ptr(), poff.data_ptr(), _ptr(gts) if gts.numel() else 0,
              _ptr(gcl) if gcl.numel() else 0, goff.data_ptr(), N, max(pc) if pc else 0, max(gc) if gc else 0, int(num_classes),
              float(iou_thresh), B, int(B * positive_fraction), int(seed), _ptr(seed_salt), int(append_gt), int(pad_background),
              _ptr(mi), _ptr(ml),
              out["sampled_idx"].data_ptr(),
              out["boxes"].data_ptr(), out["classes"].data_ptr(), out["gt_boxes"].data_ptr(), out["counts"].data_ptr(),
              _stream())
    if want_labels:
        out["matched_idx"], out["matched_label"] = mi, ml
    return out


def l2_normalize_rows(src, scale=1.0, eps=1e-12):
    """(rows, cols) fp32|bf16 -> bf16, each row scaled to `scale` / max(||row||, eps): the cosine form of the prototype
    logits (my_module.py:461-469 `sim_matrix`), as the K-contiguous operand of the logits GEMM (row stride rup8(cols))."""
    _require_cuda(src)
    rows, cols = src.shape
    assert src.stride(1) == 1 and src.dtype in (torch.float32, torch.bfloat16)
    ld = (cols + 7) // 8 * 8
    dst = torch.zeros((rows, ld), dtype=torch.bfloat16, device=src.device) if ld != cols else \
        torch.empty((rows, ld), dtype=torch.bfloat16, device=src.device)
    _lib.call("b200_l2_normalize_rows", src.data_ptr(), _dt(src), src.stride(0), dst.data_ptr(), ld, rows, cols, float(eps),
              float(scale), _stream())
    return dst[:, :cols]


class TextFusionWeights:
    """bf16 copies of the attention / predictor weights laid out for the GEMM kernel, plus the projected text
    keys/values.  Rebuilt whenever the parameters' versions change (constant at inference — the reference
    recomputes the text-side projections every forward, attentive_modules.py:274-277)."""

    def __init__(self):
        self.key = None
        self.w = {}

    @staticmethod
    def _version(params):
        return (PARAM_GENERATION[0],) + tuple((p.data_ptr(), p._version) for p in params)

    def refresh(self, named, text_parts):
        """`text_parts`: the persistent tensors the text matrix is concatenated from (class embeddings, bg row);
        keying on them — not on the freshly concatenated matrix — is what makes the cache hit."""
        params = [named[k] for k in sorted(named)] + list(text_parts)
        key = self._version(params)
        if key == self.key:
            return self.w
        text_feat = torch.cat(list(text_parts), dim=0)
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        w = {}
        for name in ("w_q.weight", "linear1.0.weight", "linear2.0.weight", "linear3.weight", "ffn.linear1.weight",
                     "ffn.linear2.weight"):
            w[name] = bf(named["attention." + name])
        for name in ("linear1.0.bias", "linear2.0.bias", "linear3.bias", "ffn.linear1.bias", "ffn.linear2.bias",
                     "ffn.norm3.weight", "ffn.norm3.bias"):
            w[name] = f32(named["attention." + name])
        # text side: tiny (K+1 rows); plain torch ops, cached
        T = text_feat.detach().float()
        kt = torch.relu(torch.nn.functional.linear(T, named["key_projection.weight"].detach().float(), named["key_projection.bias"].detach().float()))
        vt = torch.relu(torch.nn.functional.linear(T, named["value_projection.weight"].detach().float(), named["value_projection.bias"].detach().float()))
        kp = torch.nn.functional.linear(kt, named["attention.w_k.weight"].detach().float())
        vp = torch.nn.functional.linear(vt, named["attention.w_v.weight"].detach().float())
        w["kp"] = torch.cat([kp, named["attention.dummy"].detach().float().reshape(1, -1)], 0).contiguous()
        w["vp"] = torch.cat([vp, torch.zeros(1, vp.shape[1], device=vp.device)], 0).contiguous()
        w["vp_bf16"] = w["vp"].to(torch.bfloat16)
        # folded query/key operand: S = (x Wq^T) Kp^T / sqrt(d) = x (Kp Wq)^T / sqrt(d); constant while weights are
        d = w["kp"].shape[1]
        w["kq"] = bf((w["kp"] @ named["attention.w_q.weight"].detach().float()) / math.sqrt(d))
        for k in named:
            if k.startswith("extra."):
                t = named[k]
                w[k] = bf(t) if (t.dim() == 2) else f32(t)
        self.key, self.w = key, w
        return w


# inference chain: attention probabilities and gate operands as GEMM epilogues (False / B200_ATTN_EPILOGUES=0: the separate
# fp32 attention kernel of round 1, kept for A/B runs and used by the training direction, which needs its fp32 operands)
ATTENTION_AS_EPILOGUES = [_os.environ.get("B200_ATTN_EPILOGUES", "1") != "0"]
GEMM2_INFERENCE_CHAIN = [_os.environ.get("B200_GEMM2_INFER", "1") != "0"]     # 0: the round-1 single-CTA GEMM for linear1-3 / FFN


def text_fusion_forward(x, w, fold_query=True, score_bias=None):
    """A1..A6 on the device.  x (R,d) fp32.  Returns (sim2stext fp32 (R,d), sim2stext bf16, attn (R,K+2),
    x_bf16 view (R,d) with row stride 2d).  fold_query: compute the attention scores as x (Kp Wq)^T (one skinny
    GEMM against the cached folded operand) instead of Q = x Wq^T followed by Q Kp^T — same algebra, 8.4 MFLOP/ROI
    less work; fold_query=False runs the reference's literal order of operations."""
    _require_cuda(x)
    x = x.float().contiguous()
    R, d = x.shape
    dev = x.device
    h = d // 2
    xcat = torch.empty((R, 2 * d), dtype=torch.bfloat16, device=dev)      # [o1 | o2 | x]  (attentive_modules.py:172-174)
    xb = cast_bf16_into(x, xcat[:, d:])
    p1 = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
    p2 = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
    if fold_query and ATTENTION_AS_EPILOGUES[0]:
        # softmax as the epilogue of the (folded) score product, gate operands as the epilogue of probabilities x values:
        # no separate attention kernel (north_star: "fused softmax GEMM epilogues")
        L = w["kq"].shape[0]
        pb = torch.empty((R, (L + 7) // 8 * 8), dtype=torch.bfloat16, device=dev)
        attn = torch.empty((R, L), dtype=torch.float32, device=dev)
        gemm2(xb, w["kq"], bias=score_bias, softmax=True, out=pb[:, :L], out_f32=attn)    # score_bias (L,): teacher's log n_c
        vpb = w.get("vp_bf16")
        if vpb is None or vpb.shape != w["vp"].shape:
            vpb = w["vp"].to(torch.bfloat16).contiguous()
        gemm2(pb[:, :L], vpb, b_mn=True, residual=xb, gate=True, out=p1, out2=p2)
    elif fold_query:
        s = gemm_bf16(xb, w["kq"], score_bias)
        attn = text_attention(None, x, None, w["vp"], p1, p2, scores=s)
    else:
        q = gemm_bf16(xb, w["w_q.weight"], out_dtype=torch.bfloat16)
        attn = text_attention(q, x, w["kp"], w["vp"], p1, p2)
    if GEMM2_INFERENCE_CHAIN[0] and d % 8 == 0 and h % 8 == 0:
        # the CTA-pair kernel of the training direction (256-wide tiles, TMA-store epilogue, programmatic dependent launch)
        gemm2(p1, w["linear1.0.weight"], bias=w["linear1.0.bias"], relu=True, out=xcat[:, :h])
        gemm2(p2, w["linear2.0.weight"], bias=w["linear2.0.bias"], relu=True, out=xcat[:, h:d])
        y = torch.empty((R, d), dtype=torch.float32, device=dev)
        yb = gemm2(xcat, w["linear3.weight"], bias=w["linear3.bias"], out_f32=y)
        hdn = gemm2(yb, w["ffn.linear1.weight"], bias=w["ffn.linear1.bias"], relu=True)
        y2 = torch.empty((R, d), dtype=torch.float32, device=dev)
        gemm2(hdn, w["ffn.linear2.weight"], bias=w["ffn.linear2.bias"], out_f32=y2, want_out=False)
    else:
        gemm_bf16(p1, w["linear1.0.weight"], w["linear1.0.bias"], relu=True, out=xcat[:, :h])
        gemm_bf16(p2, w["linear2.0.weight"], w["linear2.0.bias"], relu=True, out=xcat[:, h:d])
        yb = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
        y = gemm_bf16(xcat, w["linear3.weight"], w["linear3.bias"], out2=yb)
        hdn = gemm_bf16(yb, w["ffn.linear1.weight"], w["ffn.linear1.bias"], relu=True, out_dtype=torch.bfloat16)
        y2 = gemm_bf16(hdn, w["ffn.linear2.weight"], w["ffn.linear2.bias"])
    z, zb = residual_layernorm(y, y2, w["ffn.norm3.weight"], w["ffn.norm3.bias"], 1e-5, relu=True)
    return z, zb, attn, xb


# ---------------------------------------------------------------------------------------------------
# A7: teacher ("language-vision") attention, forward direction — attentive_modules.py:297-487
# ---------------------------------------------------------------------------------------------------
def gather_rows_bf16(table, idx, dst, relu=False):
    """dst (rows, cols) bf16 view <- bf16(act(table[idx])) ; table (T, cols) fp32, idx (rows,) int64."""
    rows, cols = dst.shape
    assert table.dtype == torch.float32 and table.stride(1) == 1 and dst.dtype == torch.bfloat16 and dst.stride(1) == 1
    assert idx.dtype == torch.int64 and idx.is_contiguous() and idx.numel() == rows and table.shape[1] == cols
    _lib.call("b200_gather_rows_bf16", table.data_ptr(), table.stride(0), table.shape[0], idx.data_ptr(), dst.data_ptr(),
              dst.stride(0), rows, cols, int(relu), _stream())
    return dst


def class_mean_rows(x, labels, num_classes):
    """(mean (C,d) fp32, counts (C,) fp32): per-class mean of the rows of x (R,d) fp32, fixed summation order."""
    _require_cuda(x, labels)
    assert x.dtype == torch.float32 and x.stride(1) == 1 and labels.dtype == torch.int64 and labels.is_contiguous()
    R, d = x.shape
    mean = torch.empty((num_classes, d), dtype=torch.float32, device=x.device)
    counts = torch.empty(num_classes, dtype=torch.float32, device=x.device)
    nbytes = _lib.lib().b200_class_mean_rows_workspace_bytes(R, d, num_classes)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _lib.call("b200_class_mean_rows", x.data_ptr(), x.stride(0), labels.data_ptr(), R, d, num_classes, mean.data_ptr(),
              counts.data_ptr(), ws.data_ptr(), nbytes, _stream())
    return mean, counts


class TeacherFusionWeights:
    """bf16 / fp32 operand copies of a teacher module (LV_attention / LV_attention_VKV) and its label-independent text
    side: class table `proj2([embed; w_bg])`, projected keys + dummy, folded query operand.  Same caching rule as
    TextFusionWeights (constant while the teacher is frozen — the student-training and test-with-GT cases)."""

    def __init__(self):
        self.key = None
        self.w = {}

    def refresh(self, named, text_parts):
        params = [named[k] for k in sorted(named)] + list(text_parts)
        key = TextFusionWeights._version(params)
        if key == self.key:
            return self.w
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        w = {}
        for name in ("linear1.0.weight", "linear2.0.weight", "linear3.weight", "ffn.linear1.weight", "ffn.linear2.weight"):
            w[name] = bf(named["attention." + name])
        for name in ("linear1.0.bias", "linear2.0.bias", "linear3.bias", "ffn.linear1.bias", "ffn.linear2.bias",
                     "ffn.norm3.weight", "ffn.norm3.bias"):
            w[name] = f32(named["attention." + name])
        w["proj_k.weight"], w["proj_k.bias"] = bf(named["proj_k.weight"]), f32(named["proj_k.bias"])
        w["w_v"] = f32(named["attention.w_v.weight"])
        # text side: K+1 rows; plain torch ops, cached
        table = torch.nn.functional.linear(torch.cat([f32(text_parts[0]), f32(named["w_bg"])], 0),
                                           f32(named["proj2.weight"]), f32(named["proj2.bias"]))
        w["table"] = table.contiguous()
        kp = torch.nn.functional.linear(torch.relu(table), f32(named["attention.w_k.weight"]))
        kp = torch.cat([kp, f32(named["attention.dummy"]).reshape(1, -1)], 0)
        d = kp.shape[1]
        w["kq"] = bf((kp @ f32(named["attention.w_q.weight"])) / math.sqrt(d))
        self.key, self.w = key, w
        return w


def teacher_attention_forward(x, labels, w, vkv):
    """LV_attention (vkv=False) / LV_attention_VKV (vkv=True) forward on the device, no autograd.

    The reference attends every ROI over one key / value per ROI of the batch — a dense (R, R+1) attention,
    2 R (R+1) d 2 FLOP (attentive_modules.py:403-437, :452-487).  Keys are relu(table[label_j]): ROIs of one class
    share their key, so with n_c ROIs of class c and scores s_ic = q_i . k_c / sqrt(d)
        softmax_j(S)_ij V_j  summed over j  ==  sum_c softmax_c(s_ic + log n_c) . mean_{j in c} V_j      (+ the dummy key)
    i.e. a (K+2)-key attention: logits get + log n_c, values are per-class means (absent classes: -inf).  What remains
    is the student's chain (ops.text_fusion_forward) with per-batch value rows; no R x R matrix is ever formed.
    x (R,d) fp32 pooled features, labels (R,) int64 in [0, K].  Returns (sim2stext fp32 (R,d), bf16 copy)."""
    from .train_ops import skinny
    _require_cuda(x, labels)
    x = x.float().contiguous()
    labels = labels.to(torch.int64).contiguous()
    R, d = x.shape
    table = w["table"]
    C = table.shape[0]
    cat = torch.empty((R, 2 * d), dtype=torch.bfloat16, device=x.device)          # [x | t]  (:417, :466)
    cast_bf16_into(x, cat[:, :d])
    gather_rows_bf16(table, labels, cat[:, d:])
    val = gemm_bf16(cat, w["proj_k.weight"], w["proj_k.bias"], relu=True)         # value = relu(proj_k([x | t]))
    vmean, counts = class_mean_rows(val, labels, C)
    vp = torch.cat([skinny("nt", vmean, w["w_v"]), vmean.new_zeros(1, d)], 0).contiguous()
    bias = torch.cat([torch.where(counts > 0, torch.log(counts.clamp(min=1.0)), counts.new_full((), -1e30)),
                      counts.new_zeros(1)]).contiguous()
    wt = dict(w)
    wt["vp"] = vp
    z, zb, _, _ = text_fusion_forward(val if vkv else x, wt, True, score_bias=bias)
    return z, zb


# ---------------------------------------------------------------------------------------------------
# A7: LV_attention_textDomination(_VKV), forward direction — attentive_modules.py:490-687
# ---------------------------------------------------------------------------------------------------
def _pad_cols(t, ld):
    """(rows, cols) -> bf16 buffer with row pitch ld (zeros in the pitch padding), viewed as (rows, cols)."""
    out = torch.zeros((t.shape[0], ld), dtype=torch.bfloat16, device=t.device)
    out[:, : t.shape[1]].copy_(t.detach())
    return out[:, : t.shape[1]]


class TextDominationWeights:
    """bf16 operand copies of an LV_attention_textDomination(_VKV) module for the CTA-pair GEMM.  The attention runs in the
    300-d GloVe space: 300 and its half 150 are not multiples of 8 elements, which TMA needs for a row pitch, so every
    300-wide operand lives in a buffer of pitch 304 and the [o1 | o2 | q] concatenation in one of pitch 608 with the three
    blocks at columns 0, 152, 304 — the GEMMs still contract over exactly 300 / 150 columns (or over 604 with zero weight
    columns under the two-column gaps), so nothing is approximated.  Cached like TextFusionWeights."""

    def __init__(self):
        self.key, self.w = None, {}

    def refresh(self, named, text_parts):
        params = [named[k] for k in sorted(named)] + list(text_parts)
        key = TextFusionWeights._version(params)
        if key == self.key:
            return self.w
        f32 = lambda t: t.detach().float().contiguous()
        a = lambda n: named["attention." + n]
        d = a("w_q.weight").shape[0]                      # 300
        h = d // 2
        dp, hp = (d + 7) // 8 * 8, (h + 7) // 8 * 8
        w = dict(d=d, h=h, dp=dp, hp=hp)
        w["proj_visual.weight"], w["proj_visual.bias"] = named["proj_visual.weight"].detach().to(torch.bfloat16).contiguous(), f32(named["proj_visual.bias"])
        w["proj2.weight"], w["proj2.bias"] = _pad_cols(named["proj2.weight"], dp), f32(named["proj2.bias"])
        pv = named["proj_value.weight"].detach()                                   # (d, 2d): [v300 | t]
        pvp = torch.zeros((d, 2 * dp), dtype=torch.bfloat16, device=pv.device)
        pvp[:, :d].copy_(pv[:, :d])
        pvp[:, dp:dp + d].copy_(pv[:, d:])
        w["proj_value.weight"], w["proj_value.bias"] = pvp[:, : dp + d], f32(named["proj_value.bias"])
        w["linear1.weight"], w["linear1.bias"] = _pad_cols(a("linear1.0.weight"), dp), f32(a("linear1.0.bias"))
        w["linear2.weight"], w["linear2.bias"] = _pad_cols(a("linear2.0.weight"), dp), f32(a("linear2.0.bias"))
        l3 = a("linear3.weight").detach()                                          # (d, 2d): [o1 | o2 | q]
        l3p = torch.zeros((d, 2 * dp), dtype=torch.bfloat16, device=l3.device)
        l3p[:, :h].copy_(l3[:, :h])
        l3p[:, hp:hp + h].copy_(l3[:, h:d])
        l3p[:, 2 * hp:2 * hp + d].copy_(l3[:, d:])
        w["linear3.weight"], w["linear3.bias"] = l3p[:, : 2 * hp + d], f32(a("linear3.bias"))
        w["ffn1.weight"], w["ffn1.bias"] = _pad_cols(a("ffn.linear1.weight"), dp), f32(a("ffn.linear1.bias"))
        f = a("ffn.linear2.weight").shape[1]                                       # d_ffn = 1024
        w["ffn2.weight"], w["ffn2.bias"] = _pad_cols(a("ffn.linear2.weight"), (f + 7) // 8 * 8), f32(a("ffn.linear2.bias"))
        w["gamma"], w["beta"] = f32(a("ffn.norm3.weight")), f32(a("ffn.norm3.bias"))
        w["w_v"] = f32(a("w_v.weight"))
        # label-independent text side: K+1 class rows (raw GloVe rows + w_bg), keys through w_k, dummy key, folded query
        table = torch.cat([f32(text_parts[0]), f32(named["w_bg"])], 0).contiguous()
        w["table"] = table
        kp = torch.cat([torch.relu(table) @ f32(a("w_k.weight")).t(), f32(a("dummy")).reshape(1, -1)], 0)
        w["kq"] = _pad_cols((kp @ f32(a("w_q.weight"))) / math.sqrt(d), dp)
        self.key, self.w = key, w
        return w


def text_domination_forward(x, labels, w, vkv):
    """LV_attention_textDomination (vkv=False) / ..._VKV (vkv=True) forward on the device, no autograd
    (attentive_modules.py:597-634, :650-687): the visual feature is projected into the 300-d text space, attends there
    over one key / value per ROI of the batch, and is projected back (proj2).  The same class-collapse identity as
    `teacher_attention_forward` turns the (R, R+1) attention into a (K+2)-key one; every product runs on `b200_gemm2`.
    x (R, C) fp32 pooled features, labels (R,) int64 in [0, K].  Returns sim2stext (R, C) fp32."""
    _require_cuda(x, labels)
    x = x.float().contiguous()
    labels = labels.to(torch.int64).contiguous()
    R, C = x.shape
    d, h, dp, hp = w["d"], w["h"], w["dp"], w["hp"]
    dev = x.device
    f32z = lambda n: torch.zeros((R, n), dtype=torch.float32, device=dev)
    bf16z = lambda n: torch.zeros((R, n), dtype=torch.bfloat16, device=dev)
    xb = torch.empty((R, C), dtype=torch.bfloat16, device=dev)
    cast_bf16_into(x, xb)
    # v300 = proj_visual(x): fp32 (query / residual operand of the attention kernel, pitch dp) + bf16 into [v300 | t]
    cat = bf16z(2 * dp)
    v300 = f32z(dp)
    gemm2(xb, w["proj_visual.weight"], bias=w["proj_visual.bias"], out=cat[:, :d], out_f32=v300[:, :d])
    gather_rows_bf16(w["table"], labels, cat[:, dp:dp + d])
    val = f32z(dp)
    valb = gemm2(cat[:, : dp + d], w["proj_value.weight"], bias=w["proj_value.bias"], relu=True, out=bf16z(dp)[:, :d], out_f32=val[:, :d])
    ncls = w["table"].shape[0]
    vmean, counts = class_mean_rows(val[:, :d].contiguous(), labels, ncls)
    vp = torch.zeros((ncls + 1, dp), dtype=torch.float32, device=dev)
    vp[:ncls, :d] = vmean @ w["w_v"].t()
    bias = torch.cat([torch.where(counts > 0, torch.log(counts.clamp(min=1.0)), counts.new_full((), -1e30)),
                      counts.new_zeros(1)]).contiguous()
    q32, qb = (val, valb) if vkv else (v300, cat[:, :d])
    L = ncls + 1
    S = torch.empty((R, L), dtype=torch.float32, device=dev)
    gemm2(qb, w["kq"], bias=bias, out_f32=S, want_out=False)                     # scores + log n_c
    p1, p2 = bf16z(dp), bf16z(dp)
    text_attention(None, q32, None, vp, p1, p2, scores=S)                         # d = dp: the pitch padding stays zero
    xcat = bf16z(2 * dp)
    gemm2(p1[:, :d], w["linear1.weight"], bias=w["linear1.bias"], relu=True, out=xcat[:, :h])
    gemm2(p2[:, :d], w["linear2.weight"], bias=w["linear2.bias"], relu=True, out=xcat[:, hp:hp + h])
    xcat[:, 2 * hp:2 * hp + d].copy_(qb)
    y = torch.empty((R, d), dtype=torch.float32, device=dev)
    yb = gemm2(xcat[:, : 2 * hp + d], w["linear3.weight"], bias=w["linear3.bias"], out=bf16z(dp)[:, :d], out_f32=y)
    f = w["ffn2.weight"].shape[1]
    hdn = gemm2(yb, w["ffn1.weight"], bias=w["ffn1.bias"], relu=True, out=bf16z((f + 7) // 8 * 8)[:, :f])
    y2 = torch.empty((R, d), dtype=torch.float32, device=dev)
    gemm2(hdn, w["ffn2.weight"], bias=w["ffn2.bias"], out_f32=y2, want_out=False)
    z, _ = residual_layernorm(y, y2, w["gamma"], w["beta"], 1e-5, relu=True, want_f32=True, want_bf16=False)
    zb = bf16z(dp)[:, :d]
    cast_bf16_into(z, zb)
    out = torch.empty((R, C), dtype=torch.float32, device=dev)
    gemm2(zb, w["proj2.weight"], bias=w["proj2.bias"], out_f32=out, want_out=False)
    return out
