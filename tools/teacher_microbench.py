#!/usr/bin/env python
"""A7 timing: teacher attention forward (LV_attention_VKV / LV_attention), frozen teacher.  The class-collapsed fused path
(ops.teacher_attention_forward: no (R, R+1) matrix, tcgen05 GEMMs + fused attention / LayerNorm kernels) beside the
dense torch expression the reference evaluates (fp32 cuBLAS, with and without TF32).  R = 1024 is the reference's
2 images x 512 ROIs per GPU, R = 4096 the bench batch.  usage: python tools/teacher_microbench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fewshotobjectdetection_imporove_via_text_feature_b200.modeling.roi_heads import teacher_modules as tm  # noqa: E402


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[4:])[len(ts[4:]) // 2]


def main():
    d, K = 2048, 20
    gen = torch.Generator().manual_seed(1)
    embed = torch.randn(K, 300, generator=gen)
    from fewshotobjectdetection_imporove_via_text_feature_b200 import config
    cfg = config.get_cfg()
    cfg.MODEL.ADDITION.NAME = "glove"
    for name in ("LV_attention_VKV", "LV_attention"):
        m = getattr(tm, name)(d, cfg=cfg, class_embed=embed).cuda().eval()
        for R in (1024, 4096):
            x = torch.relu(torch.randn(R, d, generator=gen)).cuda()
            labels = torch.randint(0, K + 1, (R,), generator=gen).cuda()

            def fused():
                with torch.no_grad():
                    return m(x, labels)

            def dense():
                with torch.enable_grad():
                    return m(x, labels)
            a = fused()[1]["sim2stext"]
            b = dense()[1]["sim2stext"].detach()
            rel = float((a - b).norm() / b.norm())
            t_f = timed(fused)
            torch.backends.cuda.matmul.allow_tf32 = False
            t_d = timed(dense)
            torch.backends.cuda.matmul.allow_tf32 = True
            t_dt = timed(dense)
            torch.backends.cuda.matmul.allow_tf32 = False
            dense_flop = 2.0 * R * d * (2 * d + 3 * d + 2 * (R + 1) + d // 2 * 2 + 2 * d + d)
            print("%-17s R=%4d: fused %.3f ms | dense torch fp32 %.3f ms, tf32 %.3f ms | rel diff %.2e | dense path %.1f GFLOP" %
                  (name, R, t_f, t_d, t_dt, rel, dense_flop / 1e9), flush=True)


if __name__ == "__main__":
    main()
