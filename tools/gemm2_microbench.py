#!/usr/bin/env python
"""CTA-pair tcgen05 GEMM (b200_gemm2) microbench on the res5 and text-fusion shapes of the bench step (R = 4096 ROIs ->
65536 pixel rows), next to the library kernels for the same products (torch.matmul / cuDNN convolution, bf16), CUDA-event
timed with an L2 flush between launches."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200 import ops  # noqa: E402


ONLY = [x for x in os.environ.get("CASES", "").split(",") if x]
ITERS = int(os.environ.get("ITERS", "10"))


def want(name):
    return not ONLY or any(x in name for x in ONLY)


def timed(fn, flush, n=None, warm=3):
    n = ITERS if n is None else n
    ts = []
    for i in range(n + warm):
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    M = 65536
    R = M // 16
    rnd = lambda *s, sc=0.5: (torch.randn(*s, device=dev) * sc).to(torch.bfloat16)
    tiles = [int(x) for x in os.environ.get("TILES", "128,256").split(",")]
    print("%-44s %9s %9s   %s" % ("case", "ms", "TF/s", "library ms (TF/s)"))
    plain = [("res5 conv1 b0   N=512  K=1024 relu", 512, 1024, {}),
             ("res5 conv1 b1/2 N=512  K=2048 relu", 512, 2048, {}),
             ("iso N=2048 K=512 bias+relu out only", 2048, 512, {"nobits": True}),
             ("iso N=2048 K=512 + bits_out", 2048, 512, {}),
             ("iso N=2048 K=512 + res (no bits)", 2048, 512, {"res": True, "nobits": True}),
             ("res5 conv3      N=2048 K=512  relu+res", 2048, 512, {"res": True}),
             ("res5 conv3+sc   N=2048 K=512+1024 relu", 2048, 512, {"k2": 1024}),
             ("res5 dgrad3     N=512  K=2048 bits", 512, 2048, {"bits": True}),
             ("res5 dgrad1     N=2048 K=512  bits+res", 2048, 512, {"bits": True, "res": True}),
             ("res5 conv3 last N=2048 K=512 mean only", 2048, 512, {"res": True, "mean": True})]
    for name, N, K, o in plain:
        if not want(name):
            continue
        a, b = rnd(M, K), rnd(N, K + o.get("k2", 0), sc=0.05)
        a2 = rnd(M, o["k2"]) if o.get("k2") else None
        bias = torch.randn(N, device=dev)
        res = rnd(M, N) if o.get("res") else None
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (M, N // 32), dtype=torch.int32, device=dev) if o.get("bits") else None
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        bo = torch.empty(M, N // 32, dtype=torch.int32, device=dev)
        rm = torch.empty(R, N, device=dev) if o.get("mean") else None
        fl = 2.0 * M * N * (K + o.get("k2", 0))
        for tn in tiles:
            ops.GEMM2_TILE_N[0] = tn
            ms = timed(lambda: ops.gemm2(a, b, a2=a2, bias=bias, residual=res, relu=not o.get("bits"), mask_bits=bits,
                                         out=None if o.get("mean") else out, want_out=not o.get("mean"),
                                         bits_out=None if (o.get("bits") or o.get("nobits")) else bo, rowmean_out=rm), flush)
            aa = a if a2 is None else torch.cat([a, a2], 1)
            lib = timed(lambda: torch.relu_(torch.addmm(bias.to(torch.bfloat16), aa, b.t())), flush) if tn == tiles[0] else float("nan")
            print("%-40s t%-3d %9.4f %9.1f   %.4f (%.1f)" % (name, tn, ms, fl / ms / 1e9, lib, fl / lib / 1e9), flush=True)
    # 3x3 convolution
    C = 512
    x = rnd(R, C, 4, 4).contiguous(memory_format=torch.channels_last)
    w = rnd(C, C, 3, 3, sc=0.05).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, device=dev)
    xa = x.permute(0, 2, 3, 1).reshape(M, C)
    wb = w.permute(0, 2, 3, 1).reshape(C, 9 * C)
    fl = 2.0 * M * C * 9 * C
    bo = torch.empty(M, C // 32, dtype=torch.int32, device=dev)
    out = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    for tn in tiles if want("conv2 3x3") else []:
        ops.GEMM2_TILE_N[0] = tn
        ms = timed(lambda: ops.gemm2(xa, wb, conv_c=C, bias=bias, relu=True, out=out, bits_out=bo), flush)
        lib = timed(lambda: torch.cudnn_convolution_relu(x, w, bias.to(torch.bfloat16), (1, 1), (1, 1), (1, 1), 1), flush)
        print("%-40s t%-3d %9.4f %9.1f   %.4f (%.1f)" % ("res5 conv2 3x3  N=512  K=4608 relu", tn, ms, fl / ms / 1e9, lib, fl / lib / 1e9), flush=True)
    # text-fusion chain (R = 4096)
    Rr = 4096
    chain = [("linear3 fwd  M=4096 N=2048 K=4096", Rr, 2048, 4096, False, False),
             ("linear1 fwd  M=4096 N=1024 K=2048", Rr, 1024, 2048, False, False),
             ("dX (b_mn)    M=4096 N=4096 K=2048", Rr, 4096, 2048, False, True),
             ("dW (a_mn,b_mn) M=2048 N=4096 K=4096", 2048, 4096, Rr, True, True),
             ("dW (a_mn,b_mn) M=1024 N=2048 K=4096", 1024, 2048, Rr, True, True)]
    for name, Mm, N, K, amn, bmn in chain:
        if not want(name):
            continue
        a = rnd(K, Mm) if amn else rnd(Mm, K)
        b = rnd(K, N, sc=0.05) if bmn else rnd(N, K, sc=0.05)
        of = torch.empty(Mm, N, device=dev)
        fl = 2.0 * Mm * N * K
        for tn in tiles:
            ops.GEMM2_TILE_N[0] = tn
            ms = timed(lambda: ops.gemm2(a, b, a_mn=amn, b_mn=bmn, out_f32=of, want_out=False), flush)
            lib = timed(lambda: torch.matmul(a.t() if amn else a, b if bmn else b.t()), flush) if tn == tiles[0] else float("nan")
            print("%-40s t%-3d %9.4f %9.1f   %.4f (%.1f)" % (name, tn, ms, fl / ms / 1e9, lib, fl / lib / 1e9), flush=True)
    ops.GEMM2_TILE_N[0] = 0


if __name__ == "__main__":
    main()
