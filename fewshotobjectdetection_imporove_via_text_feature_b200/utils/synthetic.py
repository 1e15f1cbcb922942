"""Synthetic inputs of the shapes SURVEY.md §8(d) prescribes (no datasets or checkpoints are reachable): shared by
bench.py, the tools/ microbenchmarks, the tests and the golden-vector generator.  Pure torch on the host."""
import numpy as np
import torch


def synth_proposals(n, h, w, gen, n_obj=4):
    """70 % RPN-like boxes + 30 % jittered copies of object boxes."""
    def rand_boxes(m):
        cx = torch.rand(m, generator=gen) * w
        cy = torch.rand(m, generator=gen) * h
        side = torch.exp(torch.rand(m, generator=gen) * (np.log(min(h, w)) - np.log(16.0)) + np.log(16.0))
        asp = torch.exp((torch.rand(m, generator=gen) * 2 - 1) * np.log(3.0))
        bw, bh = side * torch.sqrt(asp), side / torch.sqrt(asp)
        b = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        b[:, 0::2] = b[:, 0::2].clamp(0, w)
        b[:, 1::2] = b[:, 1::2].clamp(0, h)
        return b
    n_j = int(0.3 * n)
    objs = rand_boxes(n_obj)
    j = objs[torch.randint(0, n_obj, (n_j,), generator=gen)] + torch.randn(n_j, 4, generator=gen) * 8.0
    j[:, 0::2] = j[:, 0::2].clamp(0, w)
    j[:, 1::2] = j[:, 1::2].clamp(0, h)
    b = torch.cat([rand_boxes(n - n_j), j], 0)
    # keep boxes non-degenerate (RPN removes empty boxes)
    b[:, 2] = torch.maximum(b[:, 2], b[:, 0] + 1.0).clamp(max=w)
    b[:, 3] = torch.maximum(b[:, 3], b[:, 1] + 1.0).clamp(max=h)
    b[:, 0] = torch.minimum(b[:, 0], b[:, 2] - 1.0).clamp(min=0)
    b[:, 1] = torch.minimum(b[:, 1], b[:, 3] - 1.0).clamp(min=0)
    return b, objs


def synth_rpn_outputs(N, level_sizes, h, w, gen, quant=0.0):
    """Decoded anchors + objectness the way an RPN head leaves them: boxes partly outside the image, clusters around a
    few objects (so NMS suppresses), logits optionally quantised (exact ties)."""
    props, logits = [], []
    for A in level_sizes:
        pb, pl = [], []
        for _ in range(N):
            b, objs = synth_proposals(A, h, w, gen, n_obj=6)
            b = b + torch.randn(A, 4, generator=gen) * 6.0 - 3.0           # un-clipped, a few inverted / outside
            l = torch.randn(A, generator=gen) * 2.0
            if quant > 0:
                l = torch.round(l / quant) * quant
            pb.append(b)
            pl.append(l)
        props.append(torch.stack(pb))
        logits.append(torch.stack(pl))
    return props, logits
