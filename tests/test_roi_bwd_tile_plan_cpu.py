"""Host-side logic of the pixel-tile ROIAlign backward (csrc/roi_align_bwd_tile.cu), no GPU: the plan-buffer size the
C ABI reports covers the worst case the header comment derives, and that worst case really bounds the number of
(tile, bin) entries of any ROI — counted here with a numpy restatement of torchvision's sampling grid
(roi_align_kernel.cpp pre_calc_for_bilinear_interpolate semantics, the call at defrcn/modeling/roi_heads/roi_heads.py:340)."""
import math

import numpy as np
import pytest

TILE, BLK_ENTRIES, BLK_BYTES, REC_BYTES = 4, 16, 576, 1024
f32 = np.float32


def _ceil_div(a, b):
    return -(-a // b)


def _align_up(x, a):
    return _ceil_div(x, a) * a


def tile_plan_bytes(N, H, W, R, PH, PW, bin_step):
    """Mirror of roi_bwd_tile_workspace_bytes."""
    pho, pwo = _ceil_div(PH, bin_step), _ceil_div(PW, bin_step)
    per_roi = _ceil_div(H + 8 * pho + 2, TILE) * _ceil_div(W + 8 * pwo + 2, TILE)
    ntiles = N * _ceil_div(H, TILE) * _ceil_div(W, TILE)
    blocks = _ceil_div(max(R, 1) * per_roi, BLK_ENTRIES) + ntiles
    return (_align_up(max(R, 1) * REC_BYTES, 256) + _align_up(ntiles * 8, 256) + 256 + _align_up(ntiles * 32 * 4, 256) +
            blocks * BLK_BYTES)


def axis_tile_bin_pairs(lo_px, hi_px, P, size, sampling_ratio, bin_step):
    """Number of (4-pixel tile, computed bin) pairs along one axis for an aligned ROI spanning [lo_px, hi_px] (map
    units): per bin the pixel range its samples touch, as the list builders see it (first lower tap .. last upper tap)."""
    start = f32(lo_px) - f32(0.5)
    length = (f32(hi_px) - f32(0.5)) - start
    bin_sz = length / f32(P)
    g = sampling_ratio if sampling_ratio > 0 else int(math.ceil(float(length / f32(P))))
    g = max(g, 0)
    pairs = 0
    for p in range(0, P, bin_step):
        first, last = None, None
        for i in range(g):
            c = (start + f32(p) * bin_sz) + (f32(i) + f32(0.5)) * bin_sz / f32(g)
            if c < -1.0 or c > size:
                continue
            c = max(c, f32(0.0))
            lo = int(c)
            if lo >= size - 1:
                lo = hi = size - 1
            else:
                hi = lo + 1
            first = lo if first is None else min(first, lo)
            last = hi if last is None else max(last, hi)
        if first is not None:
            pairs += last // TILE - first // TILE + 1
    return pairs


@pytest.mark.parametrize("H,W", [(38, 50), (50, 84), (7, 5), (256, 256)])
@pytest.mark.parametrize("bin_step,sr", [(1, 0), (2, 0), (1, 2), (1, 1)])
def test_tile_entries_of_any_roi_stay_below_the_capacity_bound(H, W, bin_step, sr):
    rng = np.random.default_rng(H * 1000 + W * 10 + bin_step + sr)
    P = 7
    po = _ceil_div(P, bin_step)
    by, bx = _ceil_div(H + 8 * po + 2, TILE), _ceil_div(W + 8 * po + 2, TILE)
    worst = 0.0
    boxes = []
    for _ in range(400):                                   # RPN-like boxes, in map units
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        h, w = np.exp(rng.uniform(np.log(0.3), np.log(1.3 * H))), np.exp(rng.uniform(np.log(0.3), np.log(1.3 * W)))
        boxes.append((cy - h / 2, cy + h / 2, cx - w / 2, cx + w / 2))
    boxes += [(0.0, float(H), 0.0, float(W)), (-10.0, H + 10.0, -10.0, W + 10.0), (3.0, 3.0, 4.0, 4.0),
              (-5.0, -2.0, -5.0, -2.0), (H - 0.5, H + 30.0, W - 0.5, W + 30.0), (0.49, 0.51, 0.49, 0.51),
              (-3 * H, 4.0 * H, -3 * W, 4.0 * W), (1.0, 1.0 + 7 * 8.0, 1.0, 1.0 + 7 * 8.0)]
    for y0, y1, x0, x1 in boxes:
        ny = axis_tile_bin_pairs(y0, y1, P, H, sr, bin_step)
        nx = axis_tile_bin_pairs(x0, x1, P, W, sr, bin_step)
        assert ny <= by and nx <= bx, (y0, y1, x0, x1, ny, by, nx, bx)
        worst = max(worst, ny * nx / (by * bx))
    assert worst > 0.02                                     # the boxes do exercise the count


def test_plan_bytes_cover_the_tile_plan():
    """b200_roi_align_bwd_plan_bytes (pure host code behind the C ABI) >= the tile plan's worst case whenever the tile
    path is eligible (C % 64 == 0), so one buffer serves whichever "roi_align_bwd_impl" builds the plan."""
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib
    L = _lib.lib()
    for (N, C, H, W, R, bs) in [(8, 1024, 38, 50, 4160, 2), (8, 1024, 38, 50, 4160, 1), (2, 64, 50, 84, 80, 1),
                                (1, 128, 7, 5, 300, 2), (16, 1024, 38, 50, 16 * 8200, 1)]:
        got = L.b200_roi_align_bwd_plan_bytes(N, C, H, W, R, 7, 7, bs)
        assert got >= tile_plan_bytes(N, H, W, R, 7, 7, bs), (N, C, H, W, R, bs, got)
        # the per-pixel lists alone (C % 64 != 0 keeps the tile path out) never need more than the common size
        assert L.b200_roi_align_bwd_plan_bytes(N, C + 8, H, W, R, 7, 7, bs) <= got
    assert L.b200_roi_align_bwd_plan_bytes(1, 1024, 300, 50, 10, 7, 7, 1) == 0        # map too tall for the byte-packed windows
    assert tile_plan_bytes(8, 38, 50, 4160, 7, 7, 2) < 200 << 20                       # bench shape: well under 200 MiB


def _axis_windows(lo_px, hi_px, P, size, bin_step):
    """Per computed bin: (first pixel, summed bilinear weights per pixel of the window) — roi_slice_prepare's tables
    (adaptive sampling grid, aligned=True), fp32 step by step like roi_geom.cuh."""
    start = f32(lo_px) - f32(0.5)
    length = (f32(hi_px) - f32(0.5)) - start
    bin_sz = length / f32(P)
    g = max(int(math.ceil(float(length / f32(P)))), 0)
    wins = []
    for p in range(0, P, bin_step):
        first, w = None, {}
        for i in range(g):
            c = (start + f32(p) * bin_sz) + (f32(i) + f32(0.5)) * bin_sz / f32(g)
            if c < -1.0 or c > size:
                continue
            c = max(c, f32(0.0))
            lo = int(c)
            if lo >= size - 1:
                lo = hi = size - 1
                c = f32(lo)
            else:
                hi = lo + 1
            l = f32(c) - f32(lo)
            first = lo if first is None else first
            w[lo] = w.get(lo, f32(0)) + (f32(1) - l)
            w[hi] = w.get(hi, f32(0)) + l
        wins.append((first, w))
    return wins, max(g, 1)


@pytest.mark.parametrize("bin_step", [1, 2])
def test_tile_gather_restatement_vs_oracle(bin_step):
    """numpy restatement of the pixel-tile backward (4 x 4 tiles; entries = (ROI, bin) pairs whose window meets the tile, in
    ROI then bin order; weight bf16(bf16(a / count) * b); fp32 accumulation; bf16 output) against the oracle's
    roi_align_bwd (pinned on torchvision) at the GPU test's bars."""
    import torch
    from oracle import oracle as O
    rng = np.random.default_rng(5 + bin_step)
    N, C, H, W, R, P, scale = 1, 8, 13, 18, 14, 7, 1 / 16
    boxes = []
    for _ in range(R):
        cy, cx = rng.uniform(0, H / scale), rng.uniform(0, W / scale)
        h, w = np.exp(rng.uniform(np.log(12), np.log(H / scale))), np.exp(rng.uniform(np.log(12), np.log(W / scale)))
        boxes.append([max(cx - w / 2, 0), max(cy - h / 2, 0), min(cx + w / 2, W / scale), min(cy + h / 2, H / scale)])
    boxes = torch.tensor(boxes, dtype=torch.float32)
    rois = O.boxes_to_rois([boxes])
    nb = _ceil_div(P, bin_step)
    g = torch.randn(R, C, nb, nb, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    gfull = torch.zeros(R, C, P, P)
    gfull[:, :, ::bin_step, ::bin_step] = g.float()
    ref = O.roi_align_bwd(gfull, rois, (N, C, H, W), scale, 0, True).numpy()

    bf = lambda v: float(torch.tensor(float(v)).to(torch.bfloat16))
    gn = g.float().numpy()
    out = np.zeros((H, W, C), np.float32)
    geo = []
    for r in range(R):
        x1, y1, x2, y2 = (f32(v) * f32(scale) for v in boxes[r].tolist())
        wy, gy = _axis_windows(y1, y2, P, H, bin_step)
        wx, gx = _axis_windows(x1, x2, P, W, bin_step)
        geo.append((wy, wx, f32(1.0) / f32(gy * gx)))
    for ty in range(0, H, TILE):
        for tx in range(0, W, TILE):
            acc = np.zeros((TILE, TILE, C), np.float32)
            for r, (wy, wx, inv) in enumerate(geo):
                for pho, (fy, ay) in enumerate(wy):
                    if fy is None or max(ay) < ty or min(ay) >= ty + TILE:
                        continue
                    for pwo, (fx, bx) in enumerate(wx):
                        if fx is None or max(bx) < tx or min(bx) >= tx + TILE:
                            continue
                        for i in range(TILE):
                            for j in range(TILE):
                                a, b = ay.get(ty + i), bx.get(tx + j)
                                if a is None or b is None:
                                    continue
                                wgt = bf(f32(bf(a * inv)) * b)
                                acc[i, j] += f32(wgt) * gn[r, :, pho, pwo]
            out[ty:ty + TILE, tx:tx + TILE] = acc[: H - ty, : W - tx]
    got = torch.tensor(out).to(torch.bfloat16).float().permute(2, 0, 1)[None].numpy()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 8e-3
    np.testing.assert_allclose(got, ref, rtol=2e-2, atol=2e-2 * np.abs(ref).max())
