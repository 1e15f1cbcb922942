"""Fine-tuning direction of the text-fused head on the C-ABI kernels (BASELINE configs[1]).

One `torch.autograd.Function` covers everything between the pooled ROI feature and the three losses:
  A1..A6 (attentive_modules.py:114-177,262-294), C1 with classifier dropout (fast_rcnn.py:403-417), and L1
  (fast_rcnn.py:222-304 `FastRCNNOutputs.losses`, roi_heads.py:1077-1081 `loss_attentive`).
Forward is the same tcgen05 GEMM chain as inference (bf16 operands, fp32 accumulate) with the activations kept for
the backward pass; backward is hand-scheduled: every `dX = dY W` and `dW = dY^T X` is the same GEMM kernel fed with
transposed bf16 copies (both of its operands are K-contiguous), ReLU backward rides on the GEMM epilogue (`mask`),
fan-in of gradients on `accumulate`, bias gradients on ordered column sums.  PyTorch supplies memory, the stream
and the autograd plumbing only; the tiny text-side projections (K+2 rows) stay in plain torch upstream of this
function and receive `dKq`, `dVp` from it.
"""
import os

import torch

from . import _lib
from . import ops as ops_mod
from ._lib import BF16, F32
from .ops import (_dt, _ptr, _require_cuda, _stream, cast_bf16_into, gemm_bf16, residual_layernorm, text_attention)


def _rup(n, m=8):
    return (n + m - 1) // m * m


def gemm_ex(a, b, bias=None, relu=False, out=None, out_dtype=torch.float32, out2=None, accumulate=False, mask=None):
    """out[M,N] (+)= act(a[M,K] @ b[N,K]^T + bias), optionally zeroed where mask <= 0 (ReLU backward)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K, (a.shape, b.shape)
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1 and tuple(out.shape) == (M, N)
    if mask is not None:
        assert mask.dtype == torch.bfloat16 and tuple(mask.shape) == (M, N) and mask.stride(1) == 1
    b32 = None if bias is None else bias.detach().float().contiguous()
    _lib.call("b200_gemm_bf16_ex", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), _ptr(b32), out.data_ptr(),
              out.stride(0), _dt(out), _ptr(out2), 0 if out2 is None else out2.stride(0), M, N, K, int(relu),
              int(accumulate), _ptr(mask), 0 if mask is None else mask.stride(0), _stream(),
              tag=(2.0 * M * N * K, (M, N, K, out.dtype == torch.bfloat16, out2 is not None, bool(relu), bool(accumulate),
                                     mask is not None, bias is not None)))
    return out


def transpose_bf16(src, pad_rows_to=None):
    """(rows, cols) fp32|bf16 with unit inner stride -> bf16 (cols[, padded], rup8(rows)); padding is zero so the
    result can be used directly as a K-contiguous GEMM operand with K = rup8(rows)."""
    rows, cols = src.shape
    assert src.stride(1) == 1
    ld = _rup(rows)
    orow = cols if pad_rows_to is None else max(cols, pad_rows_to)
    if ld != rows or orow != cols:
        dst = torch.zeros((orow, ld), dtype=torch.bfloat16, device=src.device)
    else:
        dst = torch.empty((orow, ld), dtype=torch.bfloat16, device=src.device)
    _lib.call("b200_transpose_bf16", src.data_ptr(), _dt(src), src.stride(0), dst.data_ptr(), ld, rows, cols, _stream())
    return dst


def colsum(src, n=None, out=None, accumulate=False):
    """Column sums of `src` (rows, cols) -> fp32 (cols,) [or (+)= into `out`], two-pass and ordered (bias gradients)."""
    rows, cols = src.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=src.device)
    assert out.dtype == torch.float32 and out.numel() == cols and out.is_contiguous()
    nbytes = _lib.lib().b200_colsum_workspace_bytes(cols)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
    _lib.call("b200_colsum", src.data_ptr(), _dt(src), src.stride(0), rows, cols, out.data_ptr(), int(accumulate), ws.data_ptr(),
              nbytes, _stream())
    return out if n is None else out[:n]


def dropout_bf16(x, p, seed, salt=None):
    """salt: optional int64 device scalar added to `seed` on the device (a step counter that survives graph replay)."""
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.call("b200_dropout_fwd", x.data_ptr(), y.data_ptr(), x.numel(), float(p), int(seed), _ptr(salt), _stream())
    return y


def layernorm_relu_dropout_bf16(y, y2, gamma, beta, eps, p, seed, salt=None):
    """bf16(dropout(relu(LayerNorm(y + y2) * gamma + beta))) in one pass — residual_layernorm + dropout_bf16 bit for bit."""
    R, d = y.shape
    out = torch.empty((R, d), dtype=torch.bfloat16, device=y.device)
    _lib.call("b200_residual_layernorm_dropout", y.data_ptr(), _ptr(y2), gamma.data_ptr(), beta.data_ptr(), float(eps), 1,
              float(p), int(seed), _ptr(salt), out.data_ptr(), R, d, _stream())
    return out


def head_losses(logits, deltas, attn, gt_classes, proposals, gt_boxes, K, weights, beta, acc_stats=None):
    """acc_stats (5,) fp32, optional: the counts of FastRCNNOutputs._log_accuracy from the same pass over the logits."""
    R = logits.shape[0]
    out = torch.empty(3, dtype=torch.float32, device=logits.device)
    agnostic = deltas.shape[1] == 4 and K != 1
    _lib.call("b200_head_losses", logits.data_ptr(), deltas.data_ptr(), _ptr(attn), gt_classes.data_ptr(),
              proposals.data_ptr(), gt_boxes.data_ptr(), R, K, 0 if attn is None else attn.shape[1], int(agnostic),
              *map(float, weights), float(beta), out.data_ptr(), _ptr(acc_stats), _stream())
    return out


def _nhwc(x):
    assert x.dim() == 4 and x.dtype == torch.bfloat16 and (x.numel() == 0 or x.permute(0, 2, 3, 1).is_contiguous()), \
        "expected a bf16 channels-last (R, C, h, w) tensor"
    return x


def spatial_mean(x):
    """x (R, C, h, w) bf16 channels-last -> (R, C) fp32 mean over (h, w)  [roi_heads.py:1109]."""
    _require_cuda(x)
    R, C, h, w = _nhwc(x).shape
    out = torch.empty((R, C), dtype=torch.float32, device=x.device)
    _lib.call("b200_spatial_mean", x.data_ptr(), out.data_ptr(), C, R, h * w, C, _stream())
    return out


def mean_bwd_relu_mask(gpooled, out):
    """Backward of `spatial_mean` fused with the ReLU backward of `out` (the tensor that was averaged)."""
    _require_cuda(gpooled, out)
    R, C, h, w = _nhwc(out).shape
    gp = gpooled.detach().float()
    if gp.stride(1) != 1 or gp.stride(0) % 4 or gp.data_ptr() % 16:
        gp = gp.contiguous()
    g = torch.empty_like(out)
    _lib.call("b200_mean_bwd_relu_mask", gp.data_ptr(), gp.stride(0), out.data_ptr(), g.data_ptr(), R, h * w, C, _stream())
    return g


def add_relu_mask(a, b=None, ref=None):
    """bf16(a + b) zeroed where ref <= 0; a, b, ref bf16 channels-last tensors of one shape (b / ref optional)."""
    _require_cuda(a)
    for t in (b, ref):
        assert t is None or (t.shape == a.shape and t.stride() == a.stride() and t.dtype == a.dtype)
    y = torch.empty_like(_nhwc(a))
    _lib.call("b200_add_relu_mask", a.data_ptr(), _ptr(b), _ptr(ref), y.data_ptr(), a.numel(), _stream())
    return y


# 1-bit ReLU masks for the res5 backward's elementwise passes (csrc/res5_elem.cu).  B200_RELU_BITS: 0 = every pass re-reads
# its activation; 1 (default) = the mask of the averaged tensor, which the spatial mean writes for free, feeds the mean
# backward; 2 = masks of all activations, packed on a side stream during the forward (measured: the packs are not hidden
# under cuDNN's kernels — forward +0.17 ms, backward -0.18 ms at R = 4096 — so this is not the default)
RELU_BITS = int(os.environ.get("B200_RELU_BITS", "1"))
_MASK_STREAMS = {}


def mask_stream(dev):
    key = (dev.type, dev.index)
    if key not in _MASK_STREAMS:
        _MASK_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _MASK_STREAMS[key]


def spatial_mean_bits(x):
    """`spatial_mean` that also returns the 1-bit ReLU mask of x (uint8, one byte per 8 consecutive elements)."""
    _require_cuda(x)
    R, C, h, w = _nhwc(x).shape
    out = torch.empty((R, C), dtype=torch.float32, device=x.device)
    bits = torch.empty(x.numel() // 8, dtype=torch.uint8, device=x.device)
    _lib.call("b200_spatial_mean_bits", x.data_ptr(), out.data_ptr(), C, bits.data_ptr(), R, h * w, C, _stream())
    return out, bits


def pack_relu_bits(x, bits):
    """bits (numel / 8,) uint8 <- 1-bit ReLU mask of the bf16 tensor x (dense in memory), on the current stream."""
    _require_cuda(x, bits)
    assert x.dtype == torch.bfloat16 and bits.dtype == torch.uint8 and bits.numel() * 8 == x.numel()
    _lib.call("b200_pack_relu_bits", x.data_ptr(), bits.data_ptr(), x.numel(), _stream())
    return bits


def mean_bwd_relu_bits(gpooled, bits, like):
    """`mean_bwd_relu_mask` reading the 1-bit mask of the averaged tensor (shape / layout of `like`)."""
    _require_cuda(gpooled, bits)
    R, C, h, w = _nhwc(like).shape
    gp = gpooled.detach().float()
    if gp.stride(1) != 1 or gp.stride(0) % 4 or gp.data_ptr() % 16:
        gp = gp.contiguous()
    g = torch.empty_like(like)
    _lib.call("b200_mean_bwd_relu_bits", gp.data_ptr(), gp.stride(0), bits.data_ptr(), g.data_ptr(), R, h * w, C, 0, _stream())
    return g


def add_relu_bits(a, b, bits, inplace=False):
    """bf16(a + b) (b optional) zeroed where the 1-bit mask says the activation was <= 0; `inplace` writes into a."""
    _require_cuda(a, bits)
    assert b is None or (b.shape == a.shape and b.stride() == a.stride() and b.dtype == a.dtype)
    assert bits.numel() * 8 == a.numel()
    y = a if inplace else torch.empty_like(_nhwc(a))
    _lib.call("b200_add_relu_bits", a.data_ptr(), _ptr(b), bits.data_ptr(), y.data_ptr(), a.numel(), _stream())
    return y


def skinny(mode, A, B, bias=None, relu=False, scale=1.0, relu_ref=None, out_bias=False, out=None, out_b=None):
    """fp32 contractions with a short side (csrc/text_side.cu).  mode 'nt': A (M,K), B (N,K) -> (M,N);
    'nn': A (M,N), B (N,K) -> (M,K); 'tn': A (M,N), B (M,K) -> (N,K) [+ column sums of A].  Taller A goes in row blocks
    of 32 (the big operand B is streamed once per block)."""
    A, B = A.float(), B.float()
    if A.stride(-1) != 1:
        A = A.contiguous()
    if B.stride(-1) != 1:
        B = B.contiguous()
    M = A.shape[0]
    dev = A.device
    code = {"nt": 0, "nn": 1, "tn": 2}[mode]
    if mode == "nt":
        N, K = B.shape[0], B.shape[1]
        oshape = (M, N)
    elif mode == "nn":
        N, K = B.shape
        oshape = (M, K)
    else:
        N, K = A.shape[1], B.shape[1]
        oshape = (N, K)
    if out is None:
        out = torch.empty(oshape, dtype=torch.float32, device=dev)
    assert out.dtype == torch.float32 and tuple(out.shape) == oshape and out.stride(-1) == 1
    ob = None
    if out_bias and mode == "tn":
        ob = out_b if out_b is not None else torch.empty(N, dtype=torch.float32, device=dev)
        assert ob.dtype == torch.float32 and ob.numel() == N and ob.is_contiguous()
    b32 = None if bias is None else bias.detach().float().contiguous()
    nbytes = 0 if mode == "tn" else _lib.lib().b200_skinny_gemm_workspace_bytes(min(M, 32), N if mode == "nt" else K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev) if nbytes else None
    for m0 in range(0, M, 32):
        mb = min(32, M - m0)
        a = A[m0:m0 + mb]
        ref = None if relu_ref is None else relu_ref[m0:m0 + mb]
        b = B[m0:m0 + mb] if mode == "tn" else B
        o = out if mode == "tn" else out[m0:m0 + mb]
        _lib.call("b200_skinny_gemm", code, a.data_ptr(), a.stride(0), _ptr(ref), 0 if ref is None else ref.stride(0),
                  b.data_ptr(), b.stride(0), _ptr(b32), int(relu), float(scale), o.data_ptr(), o.stride(0), _ptr(ob), mb, N, K,
                  int(mode == "tn" and m0 > 0), _ptr(ws), nbytes, _stream(), launches=1 if mode == "tn" else 2)
    return (out, ob) if (out_bias and mode == "tn") else out


def _claim_sinks(owners):
    """Deferred parameter gradients are written (not accumulated) into the optimizer's flat gradient buffer: a parameter that
    reaches a fused backward twice between two `zero_grad` calls (used twice in the graph, or micro-batch accumulation)
    would silently lose the first contribution — refuse instead."""
    for p in owners or ():
        opt = getattr(p, "_b200_opt", None)
        if opt is None:
            continue
        if getattr(p, "_b200_claim", -1) == opt.epoch:
            raise RuntimeError("FlatSGD(direct_grads=True): a parameter received a second fused gradient write in one step; "
                               "use direct_grads=False for gradient accumulation / shared parameters")
        p._b200_claim = opt.epoch


class _TextSide(torch.autograd.Function):
    """(T, key_projection, value_projection, w_k, w_v, dummy, w_q) -> (Kq (L,d) = [w_k(relu(key_proj T)); dummy] Wq / sqrt(d),
    Vp (L,d) = [w_v(relu(value_proj T)); 0]) — attentive_modules.py:274-277,125-135 plus the folded query operand."""

    @staticmethod
    def forward(ctx, T, Wkp, bkp, Wvp, bvp, Wk, Wv, dummy, Wq):
        _require_cuda(T, Wkp, Wvp, Wk, Wv, dummy, Wq)
        params = (Wkp, bkp, Wvp, bvp, Wk, Wv, dummy, Wq)
        sinks = [getattr(p, "_b200_grad_sink", None) for p in params]        # see _FusedHeadTrain.forward
        ctx.sinks = sinks if all(s is not None and s.dtype == torch.float32 and s.is_contiguous() and s.shape == p.shape and
                                 s.data_ptr() % 16 == 0 for s, p in zip(sinks, params)) else None
        ctx.sink_owners = params if ctx.sinks is not None else None
        f = lambda t: t.detach().float().contiguous()
        T, Wkp, bkp, Wvp, bvp, Wk, Wv, Wq = map(f, (T, Wkp, bkp, Wvp, bvp, Wk, Wv, Wq))
        d = Wq.shape[0]
        kt = skinny("nt", T, Wkp, bkp, relu=True)
        vt = skinny("nt", T, Wvp, bvp, relu=True)
        kp = torch.cat([skinny("nt", kt, Wk), f(dummy).reshape(1, -1)], 0)
        vp = torch.cat([skinny("nt", vt, Wv), kt.new_zeros(1, d)], 0)
        kq = skinny("nn", kp, Wq, scale=1.0 / float(d) ** 0.5)
        ctx.save_for_backward(T, kt, vt, kp, Wk, Wv, Wq)
        return kq, vp

    @staticmethod
    def backward(ctx, dkq, dvp):
        T, kt, vt, kp, Wk, Wv, Wq = ctx.saved_tensors
        d = Wq.shape[0]
        s = 1.0 / float(d) ** 0.5
        cur = torch.cuda.current_stream()
        for g in (dkq, dvp):           # produced on another stream by _FusedHeadTrain.backward (deferred mode)
            ev = _take_ready_event(g)
            if ev is not None:
                cur.wait_event(ev)
                g.record_stream(cur)
        _claim_sinks(ctx.sink_owners)
        sk = dict(zip(("Wkp", "bkp", "Wvp", "bvp", "Wk", "Wv", "dummy", "Wq"), ctx.sinks or [None] * 8))
        dkq = (dkq.float() * s).contiguous()
        dvp = dvp.float().contiguous()
        dkp = skinny("nt", dkq, Wq)                               # dKp = dKq Wq^T
        dWq = skinny("tn", kp, dkq, out=sk["Wq"])                 # dWq[i][j] = sum_l Kp[l][i] dKq[l][j]
        if ctx.sinks is not None:
            sk["dummy"].copy_(dkp[-1:].reshape(sk["dummy"].shape))
            ddummy = None
        else:
            ddummy = dkp[-1:].clone()
        dkp0, dvp0 = dkp[:-1], dvp[:-1]
        dWk = skinny("tn", dkp0, kt, out=sk["Wk"])
        dWv = skinny("tn", dvp0, vt, out=sk["Wv"])
        dkt = skinny("nn", dkp0, Wk)                              # before the ReLU mask (applied as relu_ref below)
        dvt = skinny("nn", dvp0, Wv)
        dWkp, dbkp = skinny("tn", dkt, T, relu_ref=kt, out_bias=True, out=sk["Wkp"], out_b=sk["bkp"])
        dWvp, dbvp = skinny("tn", dvt, T, relu_ref=vt, out_bias=True, out=sk["Wvp"], out_b=sk["bvp"])
        if ctx.sinks is not None:      # gradients written in place on this stream: the optimizer joins it (sync_grads)
            done = torch.cuda.Event()
            done.record(cur)
            PENDING_GRAD_EVENTS.append((done, [dkq, dvp]))
            return (None,) * 9
        return None, dWkp, dbkp, dWvp, dbvp, dWk, dWv, ddummy, dWq


def text_side(att):
    """att: SematicProposalAttention -> (Kq, Vp) with autograd through the hand-written fp32 kernels."""
    sa = att.attention
    T = att.forward_language_model()["text_feat"]
    return _TextSide.apply(T, att.key_projection.weight, att.key_projection.bias, att.value_projection.weight,
                           att.value_projection.bias, sa.w_k.weight, sa.w_v.weight, sa.dummy, sa.w_q.weight)


_TEXT_STREAMS = {}


def text_side_async(att, after=None):
    """`text_side` on a side stream: the (K+2)-row text side depends only on parameters, so its forward can run under
    the res5 kernels of the same step and — autograd replays a node on the stream its forward ran on — its backward
    under the res5 / ROIAlign backward.  `after`: an event recorded on the current stream once the parameters hold this
    step's values (e.g. at the top of the head's forward); without it the side stream waits for everything enqueued on
    the current stream so far.  Returns (kq, vp, event); the consumer waits on the event."""
    dev = att.attention.w_q.weight.device
    key = (dev.type, dev.index)
    if key not in _TEXT_STREAMS:
        _TEXT_STREAMS[key] = torch.cuda.Stream(device=dev)
    main, side = torch.cuda.current_stream(), _TEXT_STREAMS[key]
    if after is not None:
        side.wait_event(after)
    else:
        side.wait_stream(main)           # parameters were last written (optimizer step) on the current stream
    with torch.cuda.stream(side):
        kq, vp = text_side(att)
        done = torch.cuda.Event()
        done.record(side)
    return kq, vp, done


class _FusedHeadTrain(torch.autograd.Function):
    """(x, kq, vp, weights..., labels) -> (losses (3,), logits (R,K+1), accuracy counts (5,) [both non-differentiable,
    for logging])."""

    @staticmethod
    def forward(ctx, x, kq, vp, W1, b1, W2, b2, W3, b3, Wf1, bf1, Wf2, bf2, gamma, beta, Wc, bc, Wb, bb,
                gt_classes, proposals, gt_boxes, K, box_weights, l1_beta, drop_p, seed, want_attn_loss, salt=None,
                teacher_logits=None, kd=None, Wo=None, bo=None, text=None):
        """Wo, bo, text given = the CrossOutput classifier (roi_heads.py:1154-1171 + fast_rcnn.py:462-476): logits =
        relu(output_projection(z)) . text^T against the (K+1, D) text prototypes instead of cls_score(dropout(z))."""
        _require_cuda(x, kq, vp, W1, W3, Wc, Wb, gt_classes, proposals, gt_boxes)
        x = x.detach().float().contiguous()
        R, d = x.shape
        h = d // 2
        dev = x.device
        def bf(t):          # the optimizer's bf16 shadow of the parameter when it keeps one (FlatSGD), else a cast
            sh = getattr(t, "_b200_bf16", None)
            if sh is not None and sh.shape == t.shape and sh.device == t.device:
                return sh
            return t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        W = dict(W1=bf(W1), W2=bf(W2), W3=bf(W3), Wf1=bf(Wf1), Wf2=bf(Wf2), Wc=bf(Wc), Wb=bf(Wb), kq=bf(kq))
        cross = Wo is not None
        if cross:
            assert drop_p == 0.0, "the fused CrossOutput classifier has no dropout on the logits"
            W["Wo"], W["T"] = bf(Wo), bf(text)
        vpf, gam, bet = f32(vp), f32(gamma), f32(beta)
        gt = gt_classes.detach().to(torch.int64).contiguous()
        props, gtb = f32(proposals), f32(gt_boxes)

        g2 = ops_mod.gemm2
        f32e = lambda n: torch.empty((R, n), dtype=torch.float32, device=dev)
        xcat = torch.empty((R, 2 * d), dtype=torch.bfloat16, device=dev)          # [o1 | o2 | x]
        xb = cast_bf16_into(x, xcat[:, d:])
        S = f32e(W["kq"].shape[0])
        g2(xb, W["kq"], out_f32=S, want_out=False)                                 # scaled scores (R, L)
        p1 = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
        p2 = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
        attn = text_attention(None, x, None, vpf, p1, p2, scores=S)
        g2(p1, W["W1"], bias=b1, relu=True, out=xcat[:, :h])
        g2(p2, W["W2"], bias=b2, relu=True, out=xcat[:, h:d])
        y = f32e(d)
        yb = g2(xcat, W["W3"], bias=b3, out_f32=y)
        hdn = g2(yb, W["Wf1"], bias=bf1, relu=True)
        y2 = f32e(d)
        g2(hdn, W["Wf2"], bias=bf2, out_f32=y2, want_out=False)
        zd = layernorm_relu_dropout_bf16(y, y2, gam, bet, 1e-5, drop_p, seed, salt)   # sim2stext -> classifier input, one pass
        deltas = f32e(W["Wb"].shape[0])
        if cross:
            av = g2(zd, W["Wo"], bias=bo, relu=True)                               # relu(output_projection(sim2stext)), bf16
            logits = f32e(W["T"].shape[0])
            g2(av, W["T"], out_f32=logits, want_out=False)                         # dot products with the text prototypes
        else:
            av = zd
            logits = f32e(W["Wc"].shape[0])
            g2(zd, W["Wc"], bias=bc, out_f32=logits, want_out=False)
        g2(xb, W["Wb"], bias=bb, out_f32=deltas, want_out=False)
        acc = torch.empty(5, dtype=torch.float32, device=dev)
        losses = head_losses(logits, deltas, attn if want_attn_loss else None, gt, props, gtb, K, box_weights, l1_beta, acc)
        # distillation (BASELINE configs[3]): a fourth loss, KL against the frozen teacher's logits (my_module.py:409-437)
        tl = None
        if teacher_logits is not None:
            tl = f32(teacher_logits)
            assert tl.shape == logits.shape and kd is not None
            kl = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.call("b200_kd_loss", logits.data_ptr(), tl.data_ptr(), gt.data_ptr(), R, logits.shape[1], int(K), float(kd[0]),
                      float(kd[1]), kl.data_ptr(), _stream())
            losses = torch.cat([losses, kl])
        ctx.kd = (tl, kd)
        # no transposed copies anywhere: the backward's dX = dY W reads W as an N-major B operand, dW = dY^T X reads dY and
        # X as M- / N-major operands (tcgen05 smem descriptors, csrc/gemm2_tcgen05.cu)
        ctx.save_for_backward(x, xcat, p1, p2, attn, vpf, yb, y, y2, hdn, zd, logits, deltas, gam, bet, gt, props, gtb,
                              *[W[k] for k in ("W1", "W2", "W3", "Wf1", "Wf2", "Wc", "Wb", "kq")],
                              *([av, W["Wo"], W["T"]] if cross else []))
        ctx.cross = cross
        ctx.meta = (K, tuple(box_weights), float(l1_beta), float(drop_p), int(seed), bool(want_attn_loss))
        ctx.salt = salt
        # Deferred weight gradients (FlatSGD(direct_grads=True)): every parameter carries a view of the optimizer's flat
        # gradient buffer; the backward then writes dW / db straight into it from the side stream and does not join that
        # stream — the optimizer does, so the parameter-gradient work overlaps the res5 / ROIAlign backward.
        params = (W1, b1, W2, b2, W3, b3, Wf1, bf1, Wf2, bf2, gamma, beta, Wc, bc, Wb, bb) + ((Wo, bo) if cross else ())
        sinks = [getattr(p, "_b200_grad_sink", None) for p in params]
        ctx.sinks = sinks if all(s is not None and s.dtype == torch.float32 and s.is_contiguous() and s.shape == p.shape and
                                 s.data_ptr() % 16 == 0 for s, p in zip(sinks, params)) else None
        ctx.sink_owners = params if ctx.sinks is not None else None
        ctx.mark_non_differentiable(logits, acc)
        ctx.set_materialize_grads(False)             # no zero-filled gradients for the two non-differentiable outputs
        return losses, logits, acc

    @staticmethod
    def backward(ctx, g_losses, _g_logits, _g_acc):
        (x, xcat, p1, p2, attn, vp, yb, y, y2, hdn, zd, logits, deltas, gam, bet, gt, props, gtb,
         W1, W2, W3, Wf1, Wf2, Wc, Wb, kq) = ctx.saved_tensors[:26]
        cross = ctx.cross
        if cross:
            av, Wo, Tb = ctx.saved_tensors[26:]
        g2 = ops_mod.gemm2
        K, box_w, l1_beta, drop_p, seed, want_attn = ctx.meta
        R, d = x.shape
        h = d // 2
        L = attn.shape[1]
        C1, C4 = logits.shape[1], deltas.shape[1]
        C1p, C4p, Lp = _rup(C1), _rup(C4), _rup(L)
        dev = x.device
        st = _stream()
        if g_losses is None:                                 # nothing downstream used the losses
            g_losses = torch.zeros(4 if ctx.kd[0] is not None else 3, device=x.device)
        g3 = g_losses.detach().float().contiguous()          # (3,) or, with the distillation loss, (4,)

        # ---- L1 backward ------------------------------------------------------------------------------------
        dlogits = torch.empty((R, C1p), dtype=torch.bfloat16, device=dev)
        ddeltas = torch.empty((R, C4p), dtype=torch.bfloat16, device=dev)
        dattn = torch.empty((R, L), dtype=torch.float32, device=dev) if want_attn else None
        agnostic = C4 == 4 and K != 1
        _lib.call("b200_head_losses_bwd", logits.data_ptr(), deltas.data_ptr(), attn.data_ptr() if want_attn else 0,
                  gt.data_ptr(), props.data_ptr(), gtb.data_ptr(), g3.data_ptr(), R, K, L, int(agnostic),
                  *map(float, box_w), float(l1_beta), dlogits.data_ptr(), C1p, ddeltas.data_ptr(), C4p, _ptr(dattn), st)
        tl, kd = ctx.kd
        if tl is not None:
            _lib.call("b200_kd_loss_bwd", logits.data_ptr(), tl.data_ptr(), gt.data_ptr(), g3[3:].data_ptr(), R, C1, int(K),
                      float(kd[0]), float(kd[1]), dlogits.data_ptr(), C1p, st)

        # Streams: the data-gradient chain (dX GEMMs, LayerNorm / attention backward) is the critical path and stays on the
        # current stream; everything that only feeds parameter gradients (operand transposes, dW GEMMs, bias column sums)
        # runs on a side stream behind events; the two text-side operand gradients (dKq, dVp) on the text stream, where
        # their consumer (_TextSide.backward) runs.
        main = torch.cuda.current_stream()
        side = _side_stream(dev)
        sinks = ctx.sinks
        deferred = sinks is not None
        _claim_sinks(ctx.sink_owners)
        sk = dict(zip(("W1", "b1", "W2", "b2", "W3", "b3", "Wf1", "bf1", "Wf2", "bf2", "gamma", "beta", "Wc", "bc", "Wb", "bb",
                       "Wo", "bo"), sinks if deferred else [None] * 18))
        out = {}

        keep = []                       # current-stream tensors read on the other streams after this call returns

        def fork(fn, *consumed, stream=side):
            e = torch.cuda.Event()
            e.record(main)
            stream.wait_event(e)
            keep.extend(consumed)
            with torch.cuda.stream(stream):
                fn()

        def dW(dy, xin, sink):           # dW = dY^T X, fp32, both operands read in place ([K = R][M] / [K = R][N])
            out_ = sink if sink is not None else torch.empty((dy.shape[1], xin.shape[1]), dtype=torch.float32, device=dev)
            g2(dy, xin, a_mn=True, b_mn=True, out_f32=out_, want_out=False)
            return out_

        xb = xcat[:, d:]
        # ---- C1: logits = zd Wc^T + bc ; deltas = xb Wb^T + bb -------------------------------------------------
        if cross:     # logits = relu(zd Wo^T + bo) T^T: through the prototypes and the ReLU, then like cls_score
            da = g2(dlogits[:, :C1], Tb, b_mn=True, mask_act=av)             # (R, D) bf16

        def side_c1():
            if cross:
                out["dWo"] = dW(da, zd, sk.get("Wo"))
                out["dbo"] = colsum(da, out=sk.get("bo"))
                out["dWc"] = out["dbc"] = None                              # cls_score is not part of this head's graph
            else:
                out["dWc"] = dW(dlogits[:, :C1], zd, sk["Wc"])
                out["dbc"] = colsum(dlogits[:, :C1], out=sk["bc"])
            out["dWb"] = dW(ddeltas[:, :C4], xb, sk["Wb"])
            out["dbb"] = colsum(ddeltas[:, :C4], out=sk["bb"])
        fork(side_c1, dlogits, ddeltas, zd, xcat, *([da] if cross else []))
        dzd = g2(da, Wo, b_mn=True) if cross else g2(dlogits[:, :C1], Wc, b_mn=True)      # (R, d) bf16
        dx = torch.empty((R, d), dtype=torch.float32, device=dev)
        g2(ddeltas[:, :C4], Wb, b_mn=True, out_f32=dx, want_out=False)    # first producer of dL/dx (fp32)
        if _DEBUG is not None:
            _DEBUG.update(dx_box=dx.clone(), dzd=dzd.clone(), dlogits=dlogits.clone())
        # ---- A5/A6 + dropout: zd = dropout(relu(LN(y + y2))) ---------------------------------------------------
        du = torch.empty((R, d), dtype=torch.float32, device=dev)
        dub = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
        dgamma = sk["gamma"] if deferred else torch.empty(d, dtype=torch.float32, device=dev)
        dbeta = sk["beta"] if deferred else torch.empty(d, dtype=torch.float32, device=dev)
        nb = _lib.lib().b200_layernorm_bwd_workspace_bytes(R, d)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.call("b200_layernorm_relu_dropout_bwd", dzd.data_ptr(), y.data_ptr(), y2.data_ptr(), gam.data_ptr(),
                  bet.data_ptr(), 1e-5, float(drop_p), int(seed), _ptr(ctx.salt), du.data_ptr(), dub.data_ptr(), 0, 0, R, d,
                  ws.data_ptr(), nb, st, launches=1)

        def side_ln():       # dgamma / dbeta from the row statistics left in `ws`: only the optimizer needs them
            _lib.call("b200_layernorm_param_grads", dzd.data_ptr(), y.data_ptr(), y2.data_ptr(), gam.data_ptr(), bet.data_ptr(),
                      float(drop_p), int(seed), _ptr(ctx.salt), dgamma.data_ptr(), dbeta.data_ptr(), R, d, ws.data_ptr(), nb,
                      _stream())
        fork(side_ln, dzd, y, y2, gam, bet, ws)
        # ---- FFN: y2 = relu(yb Wf1^T + bf1) Wf2^T + bf2 -------------------------------------------------------
        def side_ffn2():
            out["dWf2"] = dW(dub, hdn, sk["Wf2"])
            out["dbf2"] = colsum(dub, out=sk["bf2"])   # the bf16 copy: `du` is overwritten in place by the dy GEMM below
        fork(side_ffn2, dub, hdn)
        dhdn = g2(dub, Wf2, b_mn=True, mask_act=hdn)                       # (R, h), ReLU backward fused

        def side_ffn1():
            out["dWf1"] = dW(dhdn, yb, sk["Wf1"])
            out["dbf1"] = colsum(dhdn, out=sk["bf1"])
        fork(side_ffn1, dhdn, yb)
        dyb = g2(dhdn, Wf1, b_mn=True, out_f32=du, accumulate=True)        # dy = du + dhdn Wf1 (in place), bf16 copy
        # ---- linear3: y = [o1 | o2 | xb] W3^T + b3 -----------------------------------------------------------
        def side_l3():
            out["dW3"] = dW(dyb, xcat, sk["W3"])
            out["db3"] = colsum(du, out=sk["b3"])      # du now holds dy (fp32)
        fork(side_l3, dyb, du, xcat)
        do12 = g2(dyb, W3[:, :d], b_mn=True, mask_act=xcat[:, :d])          # [do1 | do2], ReLU backward fused
        g2(dyb, W3[:, d:], b_mn=True, out_f32=dx, accumulate=True, want_out=False)
        if _DEBUG is not None:
            _DEBUG.update(dx_l3=dx.clone(), dyb=dyb.clone(), do12=do12.clone(), dub=dub.clone(), dhdn=dhdn.clone(), du=du.clone())
        # ---- linear1 / linear2: o1 = relu(P1 W1^T + b1), o2 = relu(P2 W2^T + b2) --------------------------------
        def side_l12():
            out["dW1"] = dW(do12[:, :h], p1, sk["W1"])
            out["dW2"] = dW(do12[:, h:], p2, sk["W2"])
            out["db1"] = colsum(do12[:, :h], out=sk["b1"])
            out["db2"] = colsum(do12[:, h:], out=sk["b2"])
        fork(side_l12, do12, p1, p2)
        dp1 = g2(do12[:, :h], W1, b_mn=True)
        dp2 = g2(do12[:, h:], W2, b_mn=True)
        # ---- A3 core: P1 = O * x, P2 = x - O, O = softmax(S) Vp ------------------------------------------------
        dO = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
        dS = torch.empty((R, Lp), dtype=torch.bfloat16, device=dev)
        _lib.call("b200_text_attention_bwd", dp1.data_ptr(), dp2.data_ptr(), dp1.stride(0), x.data_ptr(), attn.data_ptr(),
                  vp.data_ptr(), _ptr(dattn), dx.data_ptr(), 1, dO.data_ptr(), dS.data_ptr(), Lp, R, d, L, st)

        if _DEBUG is not None:
            _DEBUG.update(dx_att=dx.clone(), dp1=dp1.clone(), dp2=dp2.clone(), dO=dO.clone(), dS=dS.clone())
        text = _TEXT_STREAMS.get((dev.type, dev.index), side)

        def side_att():
            attn_b = torch.empty((R, Lp), dtype=torch.bfloat16, device=dev)[:, :L]
            attn_b.copy_(attn)                                                # (R, L) probabilities as a bf16 operand
            out["dvp"] = dW(attn_b, dO, None)                                 # (L, d)
            out["dkq"] = dW(dS[:, :L], xb, None)
        fork(side_att, attn, dO, dS, xcat, stream=text)
        # ---- scores: S = xb Kq^T ----------------------------------------------------------------------------------
        g2(dS[:, :L], kq, b_mn=True, out_f32=dx, accumulate=True, want_out=False)
        done = torch.cuda.Event()
        done.record(side)
        tdone = torch.cuda.Event()
        tdone.record(text)
        if deferred:
            # The parameter gradients are awaited by the optimizer (FlatSGD.sync_grads), dKq / dVp by their consumer
            # (_TextSide.backward looks the event up by the gradient's address), not by this stream.
            # Tensors those streams still read are kept referenced next to the event (dropped in sync_grads, once the
            # current stream has waited on it) rather than handed to record_stream: blocks with pending cross-stream
            # uses make the caching allocator cudaMalloc afresh when the next step asks for the same sizes.
            PENDING_GRAD_EVENTS.append((done, keep))
            for g in (out["dkq"], out["dvp"]):
                _set_ready_event(g, tdone)
            return (dx, out["dkq"], out["dvp"]) + (None,) * 16 + (None,) * 15
        main.wait_event(done)
        main.wait_event(tdone)
        dkq, dvp, dW1, dW2, dW3, dWf1, dWf2, dWc, dWb = (out[k] for k in ("dkq", "dvp", "dW1", "dW2", "dW3", "dWf1", "dWf2", "dWc", "dWb"))
        return (dx, dkq, dvp, dW1, out["db1"], dW2, out["db2"], dW3, out["db3"], dWf1, out["dbf1"], dWf2, out["dbf2"], dgamma,
                dbeta, dWc, out["dbc"], dWb, out["dbb"]) + (None,) * 12 + (out.get("dWo"), out.get("dbo"), None)


_DEBUG = None                # tests / tools: a dict that receives clones of intermediate tensors of the fused backward
_COMM_STREAMS = {}
_READY_EVENTS = {}           # (storage address, version-free) -> (weakref to the gradient, event): see _set_ready_event


def _set_ready_event(g, ev):
    """Attach "readable after `ev`" to a gradient produced on another stream.  The event rides on the tensor object
    (autograd hands the same object to the consumer node when nothing accumulates into it); the address-keyed table is the
    fall-back for the case where autograd re-wraps the storage, and it is cleared every step (FlatSGD.sync_grads)."""
    g._b200_ready = ev
    _READY_EVENTS[g.data_ptr()] = ev


def _take_ready_event(g):
    ev = getattr(g, "_b200_ready", None)
    ev2 = _READY_EVENTS.pop(g.data_ptr(), None)
    return ev if ev is not None else ev2
PENDING_GRAD_EVENTS = []     # side-stream completion events of deferred parameter gradients (drained by FlatSGD.sync_grads)
_SIDE = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


class _SplitLosses(torch.autograd.Function):
    """(n,) loss tensor -> n scalars.  Indexing the tensor instead costs autograd a zero fill and a copy per loss plus the adds
    that merge them (eight 2-us launches in the middle of the fine-tune step); here the backward is one `stack`."""

    @staticmethod
    def forward(ctx, losses):
        ctx.meta = (losses.dtype, losses.device)
        return tuple(losses.unbind(0))

    @staticmethod
    def backward(ctx, *gs):
        dt, dev = ctx.meta
        zero = None
        parts = []
        for g in gs:
            if g is None:
                if zero is None:
                    zero = torch.zeros((), dtype=dt, device=dev)
                g = zero
            parts.append(g.reshape(()))
        return torch.stack(parts)


def split_losses(losses):
    return _SplitLosses.apply(losses)


def fused_head_train(x, kq, vp, att, predictor, gt_classes, proposals, gt_boxes, K, box_weights, l1_beta, drop_p, seed,
                     want_attn_loss=True, salt=None, teacher_logits=None, kd=None, cross=None):
    """att: SingleHeadSiameseAttention (parameters linear1/2/3, ffn.*), predictor: FastRCNNOutputLayers.
    teacher_logits (R, K+1) + kd = (temperature, alpha): adds the distillation loss as a fourth entry of `losses`.
    cross = (output_projection module, text prototypes (K+1, D)): the CrossOutput classifier instead of cls_score."""
    return _FusedHeadTrain.apply(
        x, kq, vp, att.linear1[0].weight, att.linear1[0].bias, att.linear2[0].weight, att.linear2[0].bias,
        att.linear3.weight, att.linear3.bias, att.ffn.linear1.weight, att.ffn.linear1.bias, att.ffn.linear2.weight,
        att.ffn.linear2.bias, att.ffn.norm3.weight, att.ffn.norm3.bias, predictor.cls_score.weight,
        predictor.cls_score.bias, predictor.bbox_pred.weight, predictor.bbox_pred.bias, gt_classes, proposals, gt_boxes,
        K, box_weights, l1_beta, drop_p, seed, want_attn_loss, salt, teacher_logits, kd,
        *((cross[0].weight, cross[0].bias, cross[1]) if cross is not None else (None, None, None)))


class FlatSGD:
    """SGD + momentum over one flat fp32 buffer that the parameters are re-pointed into (one kernel per step;
    torch.optim.SGD semantics as configured by defrcn/solver/build.py: momentum 0.9, weight decay, no dampening).
    Every parameter starts on a 256-byte boundary of the buffer.  direct_grads=True additionally hands each parameter
    its slice of the flat gradient buffer as `_b200_grad_sink`: `_FusedHeadTrain.backward` then writes parameter
    gradients there itself, from its side stream, and `sync_grads` (called by `step`, and by the caller before a
    gradient all-reduce) is where that stream is joined.  Requires every such parameter to be used once per step."""

    def __init__(self, params, lr, momentum=0.9, weight_decay=0.0, direct_grads=False, bf16_shadow=None):
        """bf16_shadow (default: on with direct_grads): keep a flat bf16 copy of the parameters, refreshed by the update
        kernel itself; every parameter carries its slice as `_b200_bf16` and the fused head reads its GEMM operands from
        there instead of casting the fp32 weights at the top of every step."""
        self.params = [p for p in params if p.requires_grad]
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n = (n + p.numel() + 63) // 64 * 64
        self.offsets = offs + [n]
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(n, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=dev) if (direct_grads if bf16_shadow is None else bf16_shadow) else None
        for p, off in zip(self.params, offs):
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            if self.shadow is not None:
                self.shadow[off:off + k].copy_(self.flat[off:off + k])
                p._b200_bf16 = self.shadow[off:off + k].view_as(p.data)
            if direct_grads:
                p._b200_grad_sink = p.grad
                p._b200_opt = self
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self._zeroed = None
        self.epoch = 0             # bumped by zero_grad; `_claim_sinks` refuses two fused writes to one sink per epoch
        self._ranges = None        # [(begin, end)] of the flat buffer that receive gradients (decided at the first step)

    def _active_ranges(self):
        """torch.optim.SGD skips parameters whose `.grad` is None (no weight decay, no momentum) — the reference's solver
        (defrcn/solver/build.py) relies on that for parameters the live head never uses (attention.query_projection,
        attention.output_projection, ...).  With one flat gradient buffer "None" is "never written": parameters whose gradient
        slice is still all-zero after the first backward are left out of every update."""
        used = torch.stack([self.grad[o:o + p.numel()].abs().max() > 0 for p, o in zip(self.params, self.offsets)]).tolist()
        ranges = []
        for p, o, u in zip(self.params, self.offsets, used):
            if not u:
                continue
            e = (o + p.numel() + 63) // 64 * 64
            if ranges and ranges[-1][1] == o:
                ranges[-1][1] = e
            else:
                ranges.append([o, e])
        self.unused = [i for i, u in enumerate(used) if not u]
        return [tuple(r) for r in ranges]

    def zero_grad(self):
        self.epoch += 1
        self.grad.zero_()
        self._zeroed = torch.cuda.Event()
        self._zeroed.record()

    def sync_grads(self):
        """Make the current stream wait for parameter gradients still being written on side streams."""
        cur = torch.cuda.current_stream()
        while PENDING_GRAD_EVENTS:
            ev, _keepalive = PENDING_GRAD_EVENTS.pop()
            cur.wait_event(ev)
        for ev in _READY_EVENTS.values():        # text side frozen: nobody consumed dKq / dVp — still join their stream
            cur.wait_event(ev)
        _READY_EVENTS.clear()

    def all_reduce_grads(self, n_late_params=0, group=None):
        """Gradient all-reduce (replaces DDP at engine/defaults.py:252-258) for one process per GPU.  The gradients of all
        but the last `n_late_params` parameters are complete once the side streams recorded in PENDING_GRAD_EVENTS are
        (the head's parameters: frozen res5 means nothing upstream of the head trains except affine_rcnn), so their
        all-reduce goes on a communication stream behind those events and runs under the res5 / ROIAlign backward that
        the current stream still has queued; only the late tail is reduced on the current stream."""
        import torch.distributed as dist
        avg = dist.get_backend(group) == "nccl"          # ReduceOp.AVG exists on NCCL only: SUM then scale elsewhere

        def reduce_(t):
            if avg:
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                t.div_(dist.get_world_size(group))
        cur = torch.cuda.current_stream()
        dev = self.grad.device
        key = (dev.type, dev.index)
        if key not in _COMM_STREAMS:
            _COMM_STREAMS[key] = torch.cuda.Stream(device=dev)
        comm = _COMM_STREAMS[key]
        split = self.offsets[len(self.params) - n_late_params] if n_late_params else self.grad.numel()
        if not PENDING_GRAD_EVENTS or n_late_params == 0:
            self.sync_grads()
            reduce_(self.grad)
            return
        if self._zeroed is not None:
            comm.wait_event(self._zeroed)            # not before this step's zero_grad
        held = []
        while PENDING_GRAD_EVENTS:
            ev, keepalive = PENDING_GRAD_EVENTS.pop()
            comm.wait_event(ev)
            held.append(keepalive)
        with torch.cuda.stream(comm):
            reduce_(self.grad[:split])
            done = torch.cuda.Event()
            done.record(comm)
        reduce_(self.grad[split:])
        cur.wait_event(done)
        del held                                     # the current stream is now behind every stream that read them

    def step(self):
        self.sync_grads()
        if self._ranges is None:
            self._ranges = self._active_ranges()             # one host read, at the first step only
        for b, e in self._ranges:
            _lib.call("b200_sgd_momentum", self.flat[b:e].data_ptr(), self.grad[b:e].data_ptr(), self.mom[b:e].data_ptr(), e - b,
                      float(self.lr), float(self.momentum), float(self.weight_decay),
                      0 if self.shadow is None else self.shadow[b:e].data_ptr(), _stream())
        ops_mod.PARAM_GENERATION[0] += 1     # the bf16 weight caches key on this (in-place kernel updates bypass _version)


class GraphedStep:
    """Capture `fn(static_inputs) -> outputs` as one CUDA graph and replay it.

    The fine-tune step is ~200 launches on five streams; enqueueing them costs the host about as long as the GPU needs
    to run them, so a busy host stalls the GPU.  `fn` must be capture-safe: fixed shapes, no host synchronisation, every
    stream it forks joined again before it returns (`FlatSGD.sync_grads` / `step`), and step-dependent randomness keyed
    on device memory (`SematicRes5ROIHeads.use_device_dropout_counter`).  `warmup` eager calls run first on a side
    stream (they execute; the capture itself only records), then every call copies the inputs into the graph's static
    buffers and replays."""

    def __init__(self, fn, example_inputs, warmup=3):
        cur = torch.cuda.current_stream()
        self.static_in = {k: v.clone() for k, v in example_inputs.items()}
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)
        self.kernel_launches = _lib.LAUNCHES - l0        # C-ABI kernels recorded into the graph = launched per replay
        torch.cuda.synchronize()

    def __call__(self, inputs):
        """Copy `inputs` into the graph's static buffers and replay.  An input that already IS its static buffer
        (`step.static_in[k]`, e.g. the target of the caller's own host-to-device upload) is not copied."""
        for k, v in inputs.items():
            dst = self.static_in[k]
            if v is not dst and (v.data_ptr() != dst.data_ptr() or v.shape != dst.shape):
                dst.copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out
