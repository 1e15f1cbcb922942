#!/usr/bin/env python
"""tcgen05 GEMM microbench on the text-fusion chain's shapes (R = 4096 ROIs): forward epilogues (bias/ReLU, bf16 or fp32 + bf16
copy) and backward ones (ReLU mask, fp32 accumulate), CUDA-event timed with an L2 flush between launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200 import ops, train_ops  # noqa: E402


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    R = 4096
    cases = [("linear1  bf16 out, bias+relu", R, 1024, 2048, dict(relu=True, bf16=True)),
             ("linear3  fp32 + bf16 copy", R, 2048, 4096, dict(d2=True)),
             ("ffn2     fp32", R, 2048, 1024, dict()),
             ("cls      fp32 N=21", R, 21, 2048, dict()),
             ("dX       bf16 out, mask", R, 1024, 2048, dict(bf16=True, mask=True)),
             ("dX       fp32 accumulate", R, 2048, 2048, dict(acc=True)),
             ("dW       fp32 M=2048 K=4096", 2048, 4096, R, dict()),
             ("dW skinny M=24 K=4096", 24, 2048, R, dict()),
             ("res5 1x1 dgrad bf16 K=512", 65536, 2048, 512, dict(bf16=True, mask=True)),
             ("res5 1x1 dgrad bf16 K=2048", 65536, 512, 2048, dict(bf16=True, mask=True))]
    for name, M, N, K, o in cases:
        a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        b = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        mask = torch.randn(M, (N + 7) // 8 * 8, device=dev).to(torch.bfloat16)[:, :N] if o.get("mask") else None
        out = torch.zeros(M, (N + 7) // 8 * 8, device=dev, dtype=torch.bfloat16 if o.get("bf16") else torch.float32)[:, :N]
        d2 = torch.empty(M, (N + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)[:, :N] if o.get("d2") else None
        ts = []
        for i in range(13):
            flush.fill_(i & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            train_ops.gemm_ex(a, b, bias, relu=bool(o.get("relu")), out=out, out2=d2, accumulate=bool(o.get("acc")), mask=mask)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print("%-32s M=%5d N=%4d K=%4d  %.4f ms  %7.1f TF/s" % (name, M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)


if __name__ == "__main__":
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib
    for generic in [int(x) for x in os.environ.get("GEMM_GENERIC", "0").split(",")]:
        _lib.set_option("gemm_generic_epilogue", generic)
        print("---- element-wise epilogue forced: %d" % generic)
        main()
    _lib.set_option("gemm_generic_epilogue", 0)
