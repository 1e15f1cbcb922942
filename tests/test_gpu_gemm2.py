"""CTA-pair tcgen05 GEMM / implicit 3x3 convolution (csrc/gemm2_tcgen05.cu, `b200_gemm2`) vs fp64 torch on the same
bf16-rounded operands.  Tolerances: fp32 outputs 1e-3 (accumulation order only); bf16 outputs one bf16 rounding (2^-8
relative, written as rtol = atol = 1e-2 like the single-CTA kernel's test); packed masks and row gates bit-exact against
the kernel's own bf16 output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import ops
    return ops


def _rand(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(torch.bfloat16)


_POS = torch.tensor([(i >> 1) | ((i & 1) << 4) for i in range(32)], dtype=torch.int64)   # column i of a word -> its bit


def _bits_of(t):
    """packed (t > 0) words of a (M, N) tensor in b200_gemm2's layout: word n / 32, bit j <-> column 2j, bit 16 + j <->
    column 2j + 1."""
    M, N = t.shape
    b = (t > 0).reshape(M, N // 32, 32).to(torch.int64)
    w = (b << _POS).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)


def _unpack(words, N):
    w = words.cpu().to(torch.int64) & 0xffffffff
    return ((w[:, :, None] >> _POS) & 1).reshape(w.shape[0], -1)[:, :N].bool()


@pytest.mark.parametrize("tile_n", [128, 256])
@pytest.mark.parametrize("M,N,K", [(1000, 384, 200), (256, 128, 64), (4096, 2048, 1024), (77, 200, 72), (513, 520, 4096)])
def test_gemm2_plain(M, N, K, tile_n):
    ops = _ops()
    gen = torch.Generator().manual_seed(M + N + K)
    a, b = _rand(gen, M, K, scale=0.5), _rand(gen, N, K, scale=0.05)
    bias = torch.randn(N, generator=gen)
    ref = a.double() @ b.double().t() + bias.double()
    ops.GEMM2_TILE_N[0] = tile_n
    try:
        out = ops.gemm2(a.cuda(), b.cuda(), bias=bias.cuda())
        torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-2, atol=1e-2)
        of = torch.full((M, N), 7.0, device="cuda")
        out2 = torch.empty(M, N + 8, dtype=torch.bfloat16, device="cuda")[:, :N]
        outr = ops.gemm2(a.cuda(), b.cuda(), bias=bias.cuda(), relu=True, out2=out2, out_f32=of)
        torch.testing.assert_close(of.cpu().double(), ref.clamp_min(0), rtol=1e-3, atol=1e-3)
        assert torch.equal(outr, out2) and torch.equal(outr.float(), of.to(torch.bfloat16).float())
        # accumulate into the fp32 output
        ops.gemm2(a.cuda(), b.cuda(), out_f32=of, accumulate=True, want_out=False)
        torch.testing.assert_close(of.cpu().double(), ref.clamp_min(0) + ref - bias.double(), rtol=1e-3, atol=2e-3)
    finally:
        ops.GEMM2_TILE_N[0] = 0


@pytest.mark.parametrize("clusters", [1, 3])
@pytest.mark.parametrize("tile_n", [128, 256])
def test_gemm2_many_tiles_per_pair(clusters, tile_n):
    """Few CTA pairs walk many tiles: stage / accumulator / staging-buffer phases wrap several times; every epilogue
    option rides along.  The result must not depend on the number of pairs."""
    ops = _ops()
    M, N, K = 1300, 768, 328
    gen = torch.Generator().manual_seed(5)
    a, b = _rand(gen, M, K, scale=0.5).cuda(), _rand(gen, N, K, scale=0.05).cuda()
    bias = torch.randn(N, generator=gen).cuda()
    res = _rand(gen, M, N).cuda()
    act = _rand(gen, M, N).cuda()
    bits = _bits_of(act.cpu()).cuda()

    def run():
        o1 = ops.gemm2(a, b, bias=bias, residual=res, relu=True, bits_out=(bo := torch.zeros(M, N // 32, dtype=torch.int32, device="cuda")))
        o2 = ops.gemm2(a, b, residual=res, mask_bits=bits)
        o3 = ops.gemm2(a, b, mask_act=act)
        return o1, bo, o2, o3
    ops.GEMM2_TILE_N[0] = tile_n
    try:
        base = run()
        ops.GEMM2_MAX_CLUSTERS[0] = clusters
        few = run()
        ops.GEMM2_GENERIC_EPILOGUE[0] = 1            # the all-features instantiation computes the same bits
        gen_epi = run()
    finally:
        ops.GEMM2_TILE_N[0], ops.GEMM2_MAX_CLUSTERS[0], ops.GEMM2_GENERIC_EPILOGUE[0] = 0, 0, 0
    for x, y, z in zip(base, few, gen_epi):
        assert torch.equal(x, y) and torch.equal(x, z)
    o1, bo, o2, o3 = (t.cpu() for t in base)
    prod = a.cpu().double() @ b.cpu().double().t()
    ref1 = (prod + bias.cpu().double() + res.cpu().double()).clamp_min(0)
    torch.testing.assert_close(o1.double(), ref1, rtol=1e-2, atol=1e-2)
    assert torch.equal(_unpack(bo, N), o1 > 0)                      # the packed mask is the sign of the stored output
    gate = act.cpu() > 0
    torch.testing.assert_close(o2.double(), (prod + res.cpu().double()) * gate, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(o3.double(), prod * gate, rtol=1e-2, atol=1e-2)
    assert bool((o2[~gate] == 0).all()) and bool((o3[~gate] == 0).all())


@pytest.mark.parametrize("split", [2, 3, 8])
@pytest.mark.parametrize("case", ["dW_skinny", "fwd_skinny_n", "plain_bf16_mask", "conv"])
def test_gemm2_split_k(case, split):
    """split-K: the K blocks of a tile shared by several CTA pairs, partials summed in slice order by whoever arrives last.
    Same epilogues, bitwise reproducible run to run, equal (to fp32 summation order) to the unsplit launch."""
    ops = _ops()
    gen = torch.Generator().manual_seed(split)
    dev = "cuda"

    def run():
        if case == "dW_skinny":        # dW = dY^T X: M = 21 (K + 1 classes), K = R
            dy, xx = _rand(gen, 1024, 24, scale=0.5).cuda()[:, :21], _rand(gen, 1024, 2048, scale=0.5).cuda()
            of = torch.full((21, 2048), 3.0, device=dev)
            ops.gemm2(dy, xx, a_mn=True, b_mn=True, out_f32=of, accumulate=True, want_out=False)
            return of, 3.0 + dy.double().t() @ xx.double()
        if case == "fwd_skinny_n":     # logits = zd Wc^T + bc: N = 21
            a, b = _rand(gen, 700, 2048, scale=0.5).cuda(), _rand(gen, 21, 2048, scale=0.05).cuda()
            bias = torch.randn(21, generator=gen).cuda()
            of = torch.empty(700, 21, device=dev)
            ops.gemm2(a, b, bias=bias, out_f32=of, want_out=False)
            return of, a.double() @ b.double().t() + bias.double()
        if case == "plain_bf16_mask":
            a, b = _rand(gen, 300, 1024, scale=0.5).cuda(), _rand(gen, 200, 1024, scale=0.05).cuda()
            act = _rand(gen, 300, 200).cuda()
            bo = torch.zeros(300, 200 // 32 + 1, dtype=torch.int32, device=dev)
            o = ops.gemm2(a, b, relu=True, mask_act=act, out2=torch.empty(300, 208, dtype=torch.bfloat16, device=dev)[:, :200])
            return o, (a.double() @ b.double().t()).clamp_min(0) * (act > 0)
        x = _rand(gen, 16 * 20, 128, scale=0.5).cuda()
        wgt = _rand(gen, 128, 9 * 128, scale=0.05).cuda()
        o = ops.gemm2(x, wgt, conv_c=128, relu=True)
        ref = F.conv2d(x.double().reshape(20, 4, 4, 128).permute(0, 3, 1, 2), wgt.double().reshape(128, 3, 3, 128).permute(0, 3, 1, 2), padding=1)
        return o, ref.clamp_min(0).permute(0, 2, 3, 1).reshape(320, 128)
    gen.manual_seed(split)
    ops.GEMM2_SPLIT_K[0] = 1
    try:
        base, ref = run()
        ops.GEMM2_SPLIT_K[0] = split
        gen.manual_seed(split)
        a1, _ = run()
        gen.manual_seed(split)
        a2, _ = run()
    finally:
        ops.GEMM2_SPLIT_K[0] = 0
    assert torch.equal(a1, a2)
    tol = 1e-3 if a1.dtype == torch.float32 else 1e-2
    torch.testing.assert_close(a1.double(), ref, rtol=tol, atol=tol)
    torch.testing.assert_close(a1.double(), base.double(), rtol=tol, atol=tol)


def test_gemm2_k_concat():
    """out = [a | a2] b^T in one accumulator (the bottleneck's conv3 + shortcut)."""
    ops = _ops()
    M, N, K1, K2 = 2048, 512, 128, 320
    gen = torch.Generator().manual_seed(11)
    a, a2, b = _rand(gen, M, K1, scale=0.5), _rand(gen, M, K2, scale=0.5), _rand(gen, N, K1 + K2, scale=0.05)
    out = ops.gemm2(a.cuda(), b.cuda(), a2=a2.cuda(), relu=True)
    ref = (torch.cat([a, a2], 1).double() @ b.double().t()).clamp_min(0)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(512, 256, 192), (2048, 4096, 1000), (300, 200, 4096)])
def test_gemm2_transposed_operands(M, N, K, a_mn, b_mn):
    """a_mn / b_mn: the operand sits transposed in memory ((K, M) / (K, N)); no transposed copy is made."""
    ops = _ops()
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("leading dimension of a transposed operand must be a multiple of 8")
    gen = torch.Generator().manual_seed(M + N + K + 2 * a_mn + b_mn)
    a, b = _rand(gen, M, K, scale=0.5), _rand(gen, N, K, scale=0.05)
    ref = a.double() @ b.double().t()
    A = a.t().contiguous().cuda() if a_mn else a.cuda()
    B = b.t().contiguous().cuda() if b_mn else b.cuda()
    of = torch.empty(M, N, device="cuda")
    ops.gemm2(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=of, want_out=False)
    torch.testing.assert_close(of.cpu().double(), ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("R,C,N", [(64, 64, 128), (100, 512, 512), (9, 128, 256)])
def test_gemm2_conv3x3(R, C, N):
    """Implicit GEMM of a 3x3 / stride 1 / padding 1 convolution over (R, 4, 4, C) NHWC: TMA's zero fill is the padding."""
    ops = _ops()
    gen = torch.Generator().manual_seed(R + C + N)
    x = _rand(gen, R, C, 4, 4, scale=0.5)                            # NCHW values
    w = _rand(gen, N, C, 3, 3, scale=0.05)
    bias = torch.randn(N, generator=gen)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).clamp_min(0)         # (R, N, 4, 4)
    xa = x.permute(0, 2, 3, 1).contiguous().reshape(R * 16, C).cuda()                    # NHWC rows
    wb = w.permute(0, 2, 3, 1).contiguous().reshape(N, 9 * C).cuda()                     # (co, ky, kx, ci)
    out = ops.gemm2(xa, wb, conv_c=C, bias=bias.cuda(), relu=True)
    got = out.cpu().reshape(R, 4, 4, N).permute(0, 3, 1, 2).double()
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)


def test_gemm2_rowmean():
    """rowmean_out: mean over the 16 pixels of each ROI of the epilogue's fp32 values (roi_heads.py:1109)."""
    ops = _ops()
    R, N, K = 300, 2048, 512
    gen = torch.Generator().manual_seed(3)
    a, b = _rand(gen, R * 16, K, scale=0.5), _rand(gen, N, K, scale=0.05)
    res = _rand(gen, R * 16, N)
    bias = torch.randn(N, generator=gen)
    ref = (a.double() @ b.double().t() + bias.double() + res.double()).clamp_min(0)
    rm = torch.empty(R, N, device="cuda")
    bo = torch.zeros(R * 16, N // 32, dtype=torch.int32, device="cuda")
    out = ops.gemm2(a.cuda(), b.cuda(), bias=bias.cuda(), residual=res.cuda(), relu=True, rowmean_out=rm, bits_out=bo)
    torch.testing.assert_close(rm.cpu().double(), ref.reshape(R, 16, N).mean(1), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-2, atol=1e-2)
    assert torch.equal(_unpack(bo, N), out.cpu() > 0)
    # without the bf16 output (the last block of a frozen res5 keeps only the mean and the mask)
    rm2 = torch.empty(R, N, device="cuda")
    ops.gemm2(a.cuda(), b.cuda(), bias=bias.cuda(), residual=res.cuda(), relu=True, rowmean_out=rm2, want_out=False)
    assert torch.equal(rm, rm2)


def test_gemm2_full_size_res5_shapes():
    """BASELINE shapes (R = 4096 ROIs): 65536 rows; checked against torch's bf16 matmul on the same GPU."""
    ops = _ops()
    M = 65536
    gen = torch.Generator(device="cuda").manual_seed(0)
    for N, K in [(512, 1024), (2048, 512)]:
        a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
        b = (torch.randn(N, K, device="cuda", generator=gen) * 0.05).to(torch.bfloat16)
        out = ops.gemm2(a, b, relu=True)
        ref = torch.relu(a.float() @ b.float().t())
        torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


def test_gemm2_rejects_bad_arguments():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import _lib
    ops = _ops()
    a = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.B200Error):
        ops.gemm2(a[:, :60].contiguous(), a[:, :60].contiguous())     # row pitch not a multiple of 8 elements
    with pytest.raises(_lib.B200Error):
        ops.gemm2(a, a, want_out=False)                               # no output
    with pytest.raises(_lib.B200Error):
        ops.gemm2(a[:, 1:57], a[:, 1:57])                             # operand not 16-byte aligned


# ---- row-wise epilogues (one lane owns one output row): VERDICT r1 item 10 ------------------------------------------------
@pytest.mark.parametrize("tile_n", [128, 256])
@pytest.mark.parametrize("M,N,K,N2", [(1000, 512, 2048, 21), (4096, 200, 512, 81), (77, 64, 72, 5)])
def test_gemm2_cosine_logits_as_two_epilogues(M, N, K, N2, tile_n):
    """relu(x W^T + b) with the squared row norms as a by-product (rowsumsq_out), then the product with the unit-norm,
    temperature-scaled text rows scaled per row by 1 / |a| (row_scale_sumsq): the cosine logits of my_module.py:449-469 with no
    normalisation pass over the activations."""
    ops = _ops()
    gen = torch.Generator().manual_seed(M + N)
    x, w = _rand(gen, M, K, scale=0.5), _rand(gen, N, K, scale=0.05)
    bias = torch.randn(N, generator=gen) * 0.1
    t = torch.randn(N2, N, generator=gen)
    tb = (t / t.norm(dim=1, keepdim=True) * 20.0).to(torch.bfloat16)
    ops.GEMM2_TILE_N[0] = tile_n
    try:
        ssq = torch.full((M, (N + 63) // 64), -1.0, device="cuda")
        a = ops.gemm2(x.cuda(), w.cuda(), bias=bias.cuda(), relu=True, rowsumsq_out=ssq)
        v = (x.double() @ w.double().t() + bias.double()).clamp_min(0)
        torch.testing.assert_close(a.cpu().double(), v, rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(ssq.sum(1).cpu().double(), (v * v).sum(1), rtol=2e-3, atol=1e-6)
        for j in range(ssq.shape[1]):
            torch.testing.assert_close(ssq[:, j].cpu().double(), (v[:, 64 * j:64 * j + 64] ** 2).sum(1), rtol=2e-3, atol=1e-6)
        logits = torch.empty(M, N2, device="cuda")
        ops.gemm2(a, tb.cuda(), row_scale_sumsq=ssq, out_f32=logits, want_out=False)
        ad = a.cpu().double()
        ref = (ad @ tb.double().t()) / ssq.sum(1).cpu().double().sqrt().clamp_min(1e-12)[:, None]
        torch.testing.assert_close(logits.cpu().double(), ref, rtol=1e-3, atol=1e-3)
        # and against the plain cosine expression (bf16 bar)
        cos = torch.nn.functional.normalize(v, dim=1) @ torch.nn.functional.normalize(t.double(), dim=1).t() * 20.0
        assert float((logits.cpu().double() - cos).norm() / cos.norm()) < 2e-2
    finally:
        ops.GEMM2_TILE_N[0] = 0


@pytest.mark.parametrize("M,L,K", [(4096, 22, 2048), (1000, 82, 512), (333, 64, 256), (200, 128, 128), (96, 65, 64), (50, 1, 64)])
def test_gemm2_softmax_epilogue(M, L, K):
    """softmax over the row as the epilogue of the score product (attentive_modules.py:45-55): one lane owns a row, rows wider
    than 64 columns merge the two column halves' (max, sum) between the two warps of the lane quarter."""
    ops = _ops()
    gen = torch.Generator().manual_seed(M + L)
    x, kq = _rand(gen, M, K, scale=0.5), _rand(gen, L, K, scale=0.3)
    bias = torch.randn(L, generator=gen)
    Lp = (L + 7) // 8 * 8 + 8
    pb = torch.full((M, Lp), 9.0, dtype=torch.bfloat16, device="cuda")
    attn = torch.full((M, L), -1.0, device="cuda")
    ops.gemm2(x.cuda(), kq.cuda(), bias=bias.cuda(), softmax=True, out=pb[:, :L], out_f32=attn)
    ref = torch.softmax(x.double() @ kq.double().t() + bias.double(), dim=1)
    torch.testing.assert_close(attn.cpu().double(), ref, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(attn.sum(1).cpu(), torch.ones(M), rtol=1e-5, atol=1e-5)
    assert torch.equal(pb[:, :L].float(), attn.to(torch.bfloat16).float())
    # TMA stores are 16-byte granular: the columns that complete the row's last 16-byte unit receive the epilogue's value for
    # columns >= N (zero probability here, which is what the next product's K padding needs); nothing beyond is touched
    assert bool((pb[:, L:(L + 7) // 8 * 8] == 0.0).all()) and bool((pb[:, (L + 7) // 8 * 8:] == 9.0).all())
    only = torch.full((M, L), -1.0, device="cuda")
    ops.gemm2(x.cuda(), kq.cuda(), bias=bias.cuda(), softmax=True, out_f32=only, want_out=False)
    assert torch.equal(only, attn)


@pytest.mark.parametrize("M,L,d", [(4096, 22, 2048), (1000, 82, 512), (77, 6, 200)])
def test_gemm2_gate_epilogue(M, L, d):
    """O = P Vp with the gate operands O * x and x - O (attentive_modules.py:166,170) as its epilogue: P (M, L) bf16 with a
    padded pitch, Vp (L, d) read N-major (no transposed copy), x the residual operand."""
    ops = _ops()
    gen = torch.Generator().manual_seed(M + L + d)
    p = torch.softmax(torch.randn(M, L, generator=gen) * 2, dim=1).to(torch.bfloat16)
    Lp = (L + 7) // 8 * 8
    pp = torch.zeros(M, Lp, dtype=torch.bfloat16)
    pp[:, :L] = p
    dp = (d + 7) // 8 * 8
    vp = _rand(gen, L, dp)[:, :d]
    x = _rand(gen, M, dp)[:, :d]
    p1 = torch.empty(M, dp, dtype=torch.bfloat16, device="cuda")[:, :d]
    p2 = torch.empty(M, dp, dtype=torch.bfloat16, device="cuda")[:, :d]
    ops.gemm2(pp.cuda()[:, :L], vp.cuda(), b_mn=True, residual=x.cuda(), gate=True, out=p1, out2=p2)
    o = p.double() @ vp.double()
    torch.testing.assert_close(p1.cpu().double(), o * x.double(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(p2.cpu().double(), x.double() - o, rtol=1e-2, atol=1e-2)


def test_gemm2_row_epilogues_reject_bad_combinations():
    ops = _ops()
    from fewshotobjectdetection_imporove_via_text_feature_b200._lib import B200Error
    a, b = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda"), torch.zeros(256, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(B200Error):
        ops.gemm2(a, b, softmax=True)                                      # N > 128
    with pytest.raises(B200Error):
        ops.gemm2(a, b[:64], gate=True)                                    # no x / second output
    with pytest.raises(B200Error):
        ops.gemm2(a, b[:64], softmax=True, relu=True)


def test_gemm2_programmatic_dependent_launch_chain():
    """Launched with programmatic stream serialisation, a GEMM's CTAs may become resident while the previous kernel of the stream
    still runs; they must not touch global memory before griddepcontrol.wait.  A chain in which every product consumes the
    previous one's output (read-after-write through the early launch) and overwrites a buffer the previous one read
    (write-after-read), repeated, must equal the plainly serialised chain bit for bit."""
    ops = _ops()
    gen = torch.Generator().manual_seed(77)
    M, D = 4096, 1024
    x0 = _rand(gen, M, D, scale=0.5).cuda()
    ws = [_rand(gen, D, D, scale=0.03).cuda() for _ in range(6)]
    bs = [torch.randn(D, generator=gen).cuda() * 0.1 for _ in range(6)]

    def chain():
        a, b = x0.clone(), torch.empty_like(x0)
        for rep in range(4):
            for w, bias in zip(ws, bs):
                ops.gemm2(a, w, bias=bias, relu=True, out=b)       # b <- f(a); the next product reads b and overwrites a
                a, b = b, a
        return a.clone()

    outs = {}
    for no_pdl in (1, 0, 0):
        ops.GEMM2_NO_PDL[0] = no_pdl
        try:
            outs.setdefault(no_pdl, []).append(chain())
        finally:
            ops.GEMM2_NO_PDL[0] = 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[0][1])
    ref = x0.double()
    for rep in range(4):
        for w, bias in zip(ws, bs):
            ref = (ref @ w.double().t() + bias.double()).clamp_min(0).to(torch.bfloat16).double()
    assert float((outs[0][0].double() - ref).norm() / ref.norm()) < 3e-2      # 24 chained bf16 roundings
