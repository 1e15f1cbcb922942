#!/usr/bin/env python
"""Text-side fp32 products (b200_skinny_gemm) alone: CUDA events, L2 flushed before every call, median of 10.
The fine-tune step runs 13 of them (5 forward, 8 backward); next to res5's persistent GEMMs each one lands on the critical path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fewshotobjectdetection_imporove_via_text_feature_b200 import train_ops  # noqa: E402


def timed(fn, flush, n=12):
    ts = []
    for i in range(n):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    M = 22
    cases = []
    for d, t in ((2048, 512), (2048, 2048)):
        A = torch.randn(M, t, generator=g).to(dev)
        W = torch.randn(d, t, generator=g).to(dev)
        G = torch.randn(M, d, generator=g).to(dev)
        b = torch.randn(d, generator=g).to(dev)
        ref = torch.randn(M, d, generator=g).to(dev)
        out = torch.empty(d, t, device=dev)
        cases += [("nt  (%d,%d) x W(%d,%d)^T + b, relu" % (M, t, d, t), lambda A=A, W=W, b=b: train_ops.skinny("nt", A, W, b, relu=True), W.numel() * 4),
                  ("nn  (%d,%d) x W(%d,%d)" % (M, d, d, t), lambda G=G, W=W: train_ops.skinny("nn", G, W, scale=0.5), W.numel() * 4),
                  ("tn  (%d,%d)^T x (%d,%d) + colsum, relu gate" % (M, d, M, t), lambda G=G, A=A, ref=ref, out=out: train_ops.skinny("tn", G, A, relu_ref=ref, out_bias=True, out=out), out.numel() * 4)]
    print("%-58s %9s %9s" % ("product", "us", "GB/s"))
    for name, fn, nbytes in cases:
        ms = timed(fn, flush)
        print("%-58s %9.1f %9.0f" % (name, 1e3 * ms, nbytes / ms / 1e6))


if __name__ == "__main__":
    main()
