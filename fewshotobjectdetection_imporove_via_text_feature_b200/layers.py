"""Host-side building blocks around the hot path (torch plumbing, no custom arithmetic):
the detectron2-v0.3 pieces the ROI head instantiates — FrozenBN bottleneck stage (res5, cuDNN),
Box2BoxTransform, Matcher, label sub-sampling, event storage.  Names and semantics follow detectron2 so
checkpoints (`roi_heads.res5.{0,1,2}.conv1.norm.weight`, ...) load unchanged."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .structures import Boxes, Instances

try:  # pragma: no cover
    from detectron2.utils.events import get_event_storage
except Exception:  # noqa: BLE001
    class _Storage:
        def __init__(self):
            self.scalars = {}

        def put_scalar(self, name, value, smoothing_hint=True):
            self.scalars[name] = float(value)

    _STORAGE = _Storage()

    def get_event_storage():
        return _STORAGE


def cat(tensors, dim=0):
    return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)


def nonzero_tuple(x):
    return (x.unsqueeze(0) if x.dim() == 0 else x).nonzero().unbind(1)


class FrozenBatchNorm2d(nn.Module):
    _version = 3

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def scale_bias(self):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        return scale, self.bias - self.running_mean * scale

    def forward(self, x):
        s, b = self.scale_bias()
        return x * s.reshape(1, -1, 1, 1).to(x.dtype) + b.reshape(1, -1, 1, 1).to(x.dtype)


def get_norm(norm, channels):
    if not norm:
        return None
    return {"FrozenBN": FrozenBatchNorm2d, "BN": nn.BatchNorm2d}[norm](channels)


class Conv2d(nn.Conv2d):
    """nn.Conv2d with an optional `norm` child and activation (detectron2.layers.Conv2d)."""

    def __init__(self, *a, norm=None, activation=None, **k):
        super().__init__(*a, **k)
        self.norm, self.activation = norm, activation

    def forward(self, x):
        x = F.conv2d(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype), self.stride,
                     self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        return x if self.activation is None else self.activation(x)

    def folded(self, dtype, channels_last):
        """(weight, bias) with a FrozenBN child folded in; used by the inference fast path."""
        w, b = self.weight.detach().float(), None if self.bias is None else self.bias.detach().float()
        if isinstance(self.norm, FrozenBatchNorm2d):
            s, sh = self.norm.scale_bias()
            w = w * s.reshape(-1, 1, 1, 1)
            b = sh if b is None else b * s + sh
        w = w.to(dtype)
        if channels_last:
            w = w.contiguous(memory_format=torch.channels_last)
        return w, None if b is None else b.to(dtype)


class _FrozenBottleneckFn(torch.autograd.Function):
    """Bottleneck block with frozen, BN-folded weights (ROI_HEADS.FREEZE_FEAT fine-tuning): the forward runs cuDNN's
    runtime-fused conv+bias+ReLU(+residual) kernels exactly like inference, the backward propagates only the data
    gradient (no weight gradients exist) — one cuDNN dgrad per convolution plus the ReLU masks.  res5 stays on
    cuDNN; this node only removes the autograd overhead of the unfused expression."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, wsc, bsc, s1, ssc, conv2_args):
        stride2, pad2, dil2, groups2 = conv2_args
        out1 = torch.cudnn_convolution_relu(x, w1, b1, s1, (0, 0), (1, 1), 1)
        out2 = torch.cudnn_convolution_relu(out1, w2, b2, stride2, pad2, dil2, groups2)
        if wsc is None:
            res, bias3 = x, b3
        else:
            res = F.conv2d(x, wsc, None, ssc)
            bias3 = b3 if bsc is None else b3 + bsc
        out = torch.cudnn_convolution_add_relu(out2, w3, res, 1.0, bias3, (1, 1), (0, 0), (1, 1), 1)
        ctx.save_for_backward(x, out1, out2, out, w1, w2, w3, wsc if wsc is not None else w1)
        ctx.meta = (s1, ssc, conv2_args, wsc is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out1, out2, out, w1, w2, w3, wsc = ctx.saved_tensors
        s1, ssc, (stride2, pad2, dil2, groups2), has_sc = ctx.meta
        relu_bwd = torch.ops.aten.threshold_backward

        def dgrad(g_out, inp, w, stride=(1, 1), pad=(0, 0), dil=(1, 1), groups=1):
            # data gradient only (output_mask): the input tensor is passed for its shape / layout, its values are unused
            return torch.ops.aten.convolution_backward(g_out, inp, w, None, list(stride), list(pad), list(dil), False, [0, 0],
                                                       groups, [True, False, False])[0]

        g = relu_bwd(g.contiguous(memory_format=torch.channels_last), out, 0)
        g2 = relu_bwd(dgrad(g, out2, w3), out2, 0)
        g1 = relu_bwd(dgrad(g2, out1, w2, stride2, pad2, dil2, groups2), out1, 0)
        gx = dgrad(g1, x, w1, s1)
        gx = gx + (dgrad(g, x, wsc, ssc) if has_sc else g)
        return (gx,) + (None,) * 11


class _FrozenRes5MeanFn(torch.autograd.Function):
    """Whole frozen res5 stage + spatial mean as one autograd node (roi_heads.py:313-344 + :1109 under
    ROI_HEADS.FREEZE_FEAT): pooled ROI map (R, C, h, w) bf16 channels-last -> (R, C_out) fp32.

    Convolutions stay on cuDNN / cuBLAS (fused conv+bias+ReLU(+residual) forward, data-gradient-only backward); the
    elementwise passes between them are the C-ABI kernels of csrc/res5_elem.cu — spatial mean, mean-backward fused
    with the last ReLU mask, residual fan-in fused with the previous block's ReLU mask — instead of one torch kernel
    per algebraic step (0.81 + 3 x 0.22 ms of a 7.8 ms step in the round-1 launch list)."""

    @staticmethod
    def forward(ctx, x, metas, *params):
        from . import train_ops
        saved = []
        y = x
        # 1-bit ReLU masks of activations whose only use in the backward is `> 0` (csrc/res5_elem.cu): the last one falls
        # out of the spatial mean for free (default); RELU_BITS = 2 also packs the others on a side stream
        last_bits = int(train_ops.RELU_BITS) >= 1 and ctx.needs_input_grad[0]
        use_bits = int(train_ops.RELU_BITS) >= 2 and ctx.needs_input_grad[0] and _nhwc_dense(x)
        cur = torch.cuda.current_stream()
        side = train_ops.mask_stream(x.device) if use_bits else None
        bits = []

        def pack_async(t):
            b = torch.empty(t.numel() // 8, dtype=torch.uint8, device=t.device)
            ev = torch.cuda.Event()
            ev.record(cur)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                train_ops.pack_relu_bits(t, b)
            return b

        for i, (s1, ssc, conv2_args, has_sc) in enumerate(metas):
            w1, b1, w2, b2, w3, b3, wsc, bsc = params[8 * i: 8 * i + 8]
            stride2, pad2, dil2, groups2 = conv2_args
            out1 = torch.cudnn_convolution_relu(y, w1, b1, s1, (0, 0), (1, 1), 1)
            use_bits = use_bits and _nhwc_dense(out1)
            m1 = pack_async(out1) if use_bits else None
            out2 = torch.cudnn_convolution_relu(out1, w2, b2, stride2, pad2, dil2, groups2)
            m2 = pack_async(out2) if use_bits and _nhwc_dense(out2) else None
            if has_sc:
                res = F.conv2d(y, wsc, None, ssc)
                bias3 = b3 if bsc is None else b3 + bsc
            else:
                res, bias3 = y, b3
            out = torch.cudnn_convolution_add_relu(out2, w3, res, 1.0, bias3, (1, 1), (0, 0), (1, 1), 1)
            my = pack_async(out) if use_bits and m2 is not None and i + 1 < len(metas) and _nhwc_dense(out) else None
            saved += [y, out1, out2, w1, w2, w3, wsc if has_sc else w1]
            bits += [m1 if m2 is not None else None, m2, my]
            y = out
        if last_bits and _nhwc_dense(y):
            pooled, mlast = train_ops.spatial_mean_bits(y)
        else:
            pooled, mlast = train_ops.spatial_mean(y), None
        if side is not None:
            cur.wait_stream(side)        # the masks (and the activations they were read from) are safe to free from here on
        ctx.save_for_backward(y, *saved)
        ctx.metas = metas
        ctx.bits = (bits, mlast)
        return pooled

    @staticmethod
    def backward(ctx, gp):
        from . import train_ops
        out_last, *saved = ctx.saved_tensors
        metas = ctx.metas
        bits, mlast = ctx.bits
        relu_bwd = torch.ops.aten.threshold_backward

        def dgrad(g_out, inp, w, stride=(1, 1), pad=(0, 0), dil=(1, 1), groups=1):
            # data gradient only (output_mask): the input tensor is passed for its shape / layout, its values are unused
            return torch.ops.aten.convolution_backward(g_out, inp, w, None, list(stride), list(pad), list(dil), False, [0, 0],
                                                       groups, [True, False, False])[0]

        def relu_back(g_raw, act, m):
            if m is not None and _nhwc_dense(g_raw):
                return train_ops.add_relu_bits(g_raw, None, m, inplace=True)
            return relu_bwd(g_raw, act, 0)

        g = train_ops.mean_bwd_relu_mask(gp, out_last) if mlast is None else train_ops.mean_bwd_relu_bits(gp, mlast, out_last)
        for i in reversed(range(len(metas))):
            s1, ssc, (stride2, pad2, dil2, groups2), has_sc = metas[i]
            x, out1, out2, w1, w2, w3, wsc = saved[7 * i: 7 * i + 7]
            m1, m2, _ = bits[3 * i: 3 * i + 3]
            mx = bits[3 * (i - 1) + 2] if i > 0 else None
            g2 = relu_back(dgrad(g, out2, w3), out2, m2)
            g1 = relu_back(dgrad(g2, out1, w2, stride2, pad2, dil2, groups2), out1, m1)
            gx = dgrad(g1, x, w1, s1)
            # fan-in of the two branches (shortcut convolution or identity); for i > 0 the block input is the previous
            # block's post-ReLU output, whose mask is applied in the same pass
            gsc = dgrad(g, x, wsc, ssc) if has_sc else g
            cl = torch.channels_last
            gx, gsc = gx.contiguous(memory_format=cl), gsc.contiguous(memory_format=cl)
            if i > 0 and mx is not None:
                g = train_ops.add_relu_bits(gx, gsc, mx)
            else:
                g = train_ops.add_relu_mask(gx, gsc, x if i > 0 else None)
        return (g, None) + (None,) * (8 * len(metas))


def _nhwc_dense(t):
    """bf16 4-d tensor whose channels-last memory is dense (flat element order == (n, h, w, c))."""
    return t.dim() == 4 and t.dtype == torch.bfloat16 and t.numel() % 8 == 0 and t.permute(0, 2, 3, 1).is_contiguous()


def frozen_res5_mean(blocks, x, prestrided=False):
    """res5 (frozen, BN folded) + mean over (h, w) -> (R, C_out) fp32 through `_FrozenRes5MeanFn`; None when the fused
    node does not apply (CPU tensors, cuDNN fused ops unavailable, non-bf16 / non-channels-last input)."""
    if not (_FUSED_CONV["ok"] and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[0] > 0 and
            x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()):
        return None
    metas, params = [], []
    for i, blk in enumerate(blocks):
        (w1, b1), (w2, b2), (w3, b3), sc, s1, ssc = blk.folded_params(x, prestrided and i == 0)
        c2 = blk.conv2
        metas.append((s1, ssc, (c2.stride, c2.padding, c2.dilation, c2.groups), sc is not None))
        params += [w1, b1, w2, b2, w3, b3, None if sc is None else sc[0], None if sc is None else sc[1]]
    return _FrozenRes5MeanFn.apply(x, tuple(metas), *params)


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, num_groups=1, norm="BN",
                 stride_in_1x1=False, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.shortcut = None
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False,
                                   norm=get_norm(norm, out_channels))
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=s1, bias=False,
                            norm=get_norm(norm, bottleneck_channels))
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=s3, padding=dilation,
                            bias=False, groups=num_groups, dilation=dilation, norm=get_norm(norm, bottleneck_channels))
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False, norm=get_norm(norm, out_channels))
        for m in (self.conv1, self.conv2, self.conv3, self.shortcut):
            if m is not None:
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._fold_key, self._fold = None, None

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        sc = x if self.shortcut is None else self.shortcut(x)
        out += sc
        return F.relu_(out)

    def reads_strided_1x1(self):
        """True when this block only ever reads every `stride`-th pixel of its input: conv1 and the shortcut are both
        1x1 convolutions carrying the stride (RESNETS.STRIDE_IN_1X1=True).  The pooler can then skip the dead bins."""
        s = (self.stride, self.stride)
        return (self.stride > 1 and self.shortcut is not None and self.conv1.kernel_size == (1, 1) and
                self.shortcut.kernel_size == (1, 1) and self.conv1.stride == s and self.shortcut.stride == s and
                self.conv1.padding == (0, 0) and self.shortcut.padding == (0, 0))

    def folded_params(self, x, prestrided=False):
        """FrozenBN folded into the conv weights, cached per (dtype, layout, parameter versions):
        ((w1, b1), (w2, b2), (w3, b3), shortcut (w, b) | None, conv1 stride, shortcut stride)."""
        s1 = (1, 1) if prestrided else self.conv1.stride
        ssc = (1, 1) if (prestrided or self.shortcut is None) else self.shortcut.stride
        convs = [self.conv1, self.conv2, self.conv3, self.shortcut]
        cl = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        key = (x.dtype, cl, x.device) + tuple((c.weight.data_ptr(), c.weight._version) for c in convs if c is not None)
        if key != self._fold_key:
            self._fold = [None if c is None else c.folded(x.dtype, cl) for c in convs]
            self._fold_key = key
        return tuple(self._fold) + (s1, ssc)

    def forward_folded(self, x, prestrided=False):
        """Frozen-weights path.  prestrided: `x` already holds only the pixels [::stride, ::stride] (see
        `reads_strided_1x1`), so conv1 and the shortcut run with stride 1 — same numbers, 1/stride^2 of the input."""
        (w1, b1), (w2, b2), (w3, b3), sc, s1, ssc = self.folded_params(x, prestrided)
        if _FUSED_CONV["ok"] and x.is_cuda and torch.is_grad_enabled() and x.requires_grad:
            return _FrozenBottleneckFn.apply(
                x, w1, b1, w2, b2, w3, b3, None if sc is None else sc[0], None if sc is None else sc[1], s1, ssc,
                (self.conv2.stride, self.conv2.padding, self.conv2.dilation, self.conv2.groups))
        if _FUSED_CONV["ok"] and x.is_cuda and not torch.is_grad_enabled():
            try:   # cuDNN runtime-fused conv + bias + ReLU (+ residual add): no separate elementwise kernels
                out = torch.cudnn_convolution_relu(x, w1, b1, s1, (0, 0), (1, 1), 1)
                out = torch.cudnn_convolution_relu(out, w2, b2, self.conv2.stride, self.conv2.padding, self.conv2.dilation, self.conv2.groups)
                # the shortcut's (folded-BN) bias rides on conv3's fused bias: a biased F.conv2d would add it with a
                # separate broadcast-add kernel over the whole (R,2048,4,4) tensor (ncu: 297 us of a 2.7 ms step)
                if sc is None:
                    res, bias3 = x, b3
                else:
                    res = F.conv2d(x, sc[0], None, ssc)
                    bias3 = b3 if sc[1] is None else b3 + sc[1]
                return torch.cudnn_convolution_add_relu(out, w3, res, 1.0, bias3, (1, 1), (0, 0), (1, 1), 1)
            except RuntimeError as e:
                # only "this cuDNN build has no runtime-fused conv+bias+ReLU for this configuration" selects the unfused
                # path; out-of-memory and asynchronous CUDA faults are real errors and must surface
                msg = str(e).lower()
                if "out of memory" in msg or "cuda error" in msg or "illegal" in msg or "launch failure" in msg:
                    raise
                import warnings
                warnings.warn("b200roi: cuDNN fused conv+bias+ReLU unavailable (%s); using the unfused library path" %
                              str(e).splitlines()[0][:160])
                _FUSED_CONV["ok"] = False
        out = F.relu_(F.conv2d(x, w1, b1, s1))
        out = F.relu_(F.conv2d(out, w2, b2, self.conv2.stride, self.conv2.padding, self.conv2.dilation, self.conv2.groups))
        out = F.conv2d(out, w3, b3)
        out += x if sc is None else F.conv2d(x, sc[0], sc[1], ssc)
        return F.relu_(out)


_FUSED_CONV = {"ok": True}


def make_stage(block_class, num_blocks, first_stride, *, in_channels, out_channels, **kwargs):
    blocks = []
    for i in range(num_blocks):
        blocks.append(block_class(in_channels=in_channels, out_channels=out_channels,
                                  stride=first_stride if i == 0 else 1, **kwargs))
        in_channels = out_channels
    return blocks


_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    """(dx,dy,dw,dh) parameterisation.  `apply_deltas` here is the torch form used on the training path and in
    tests; inference decodes inside the compaction kernel (csrc/detect_post.cu)."""

    def __init__(self, weights, scale_clamp=_SCALE_CLAMP):
        self.weights, self.scale_clamp = tuple(weights), scale_clamp

    def get_deltas(self, src, dst):
        sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
        sx, sy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
        tw, th = dst[:, 2] - dst[:, 0], dst[:, 3] - dst[:, 1]
        tx, ty = dst[:, 0] + 0.5 * tw, dst[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        return torch.stack((wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)), dim=1)

    def apply_deltas(self, deltas, boxes):
        boxes = boxes.to(deltas.dtype)
        w, h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        cx, cy = boxes[:, 0] + 0.5 * w, boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
        dw = torch.clamp(deltas[:, 2::4] / ww, max=self.scale_clamp)
        dh = torch.clamp(deltas[:, 3::4] / wh, max=self.scale_clamp)
        pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
        pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
        out = torch.zeros_like(deltas)
        out[:, 0::4], out[:, 1::4] = pcx - 0.5 * pw, pcy - 0.5 * ph
        out[:, 2::4], out[:, 3::4] = pcx + 0.5 * pw, pcy + 0.5 * ph
        return out


class Matcher:
    """argmax-IoU assignment with thresholded labels (detectron2.modeling.matcher.Matcher, no low-quality path)."""

    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        assert not allow_low_quality_matches, "ROI heads never enable low-quality matches"
        self.thresholds = [-float("inf")] + list(thresholds) + [float("inf")]
        self.labels = list(labels)

    def __call__(self, iou):  # iou: (M gt, N proposals)
        if iou.numel() == 0:
            return (iou.new_zeros((iou.size(1),), dtype=torch.int64),
                    iou.new_full((iou.size(1),), self.labels[0], dtype=torch.int8))
        vals, idx = iou.max(dim=0)
        lab = idx.new_full(idx.size(), 1, dtype=torch.int8)
        for l, lo, hi in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            lab[(vals >= lo) & (vals < hi)] = l
        return idx, lab


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    pos = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    neg = nonzero_tuple(labels == bg_label)[0]
    n_pos = min(pos.numel(), int(num_samples * positive_fraction))
    n_neg = min(neg.numel(), num_samples - n_pos)
    p1 = torch.randperm(pos.numel(), device=pos.device)[:n_pos]
    p2 = torch.randperm(neg.numel(), device=neg.device)[:n_neg]
    return pos[p1], neg[p2]


def add_ground_truth_to_proposals(gt_boxes, proposals):
    out = []
    logit = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    for gt, p in zip(gt_boxes, proposals):
        g = Instances(p.image_size)
        g.proposal_boxes = gt
        g.objectness_logits = logit * torch.ones(len(gt), device=gt.tensor.device)
        out.append(Instances.cat([p, g]))
    return out


def smooth_l1_loss(input, target, beta, reduction="none"):
    n = torch.abs(input - target)
    loss = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "sum":
        return loss.sum()
    if reduction == "mean":
        return loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    return loss
